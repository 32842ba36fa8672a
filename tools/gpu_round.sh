#!/bin/bash
# One GPU-box session: tests, smoke, bench, ncu launch list + full capture. Outputs under gpurun_out/.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench hand"; timeout 600 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_hand.json
echo "=== bench elbow"; timeout 600 python bench.py --env myoElbowPose1D6MRandom-v0 --steps 200 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_elbow.json
if [ "$1" == "ncu" ]; then
echo "=== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log
echo "=== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:myo_env_kernel -s 4 -c 2 -f -o gpurun_out/prof_hand python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
fi
