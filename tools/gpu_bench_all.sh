#!/bin/bash
# One GPU-box session: tests, smoke, the BASELINE configs on one GPU, the reference arm, ncu launch lists + full capture of the hand step.
# Outputs under gpurun_out/ (copy what should be judged into profiles/).
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-300
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 150 python bench.py --steps 100 --warmup 5 | tee gpurun_out/bench_hand.json | cut -c1-400
timeout 60 python bench.py --env myoElbowPose1D6MRandom-v0 --steps 300 --warmup 5 --no-cpu-baseline | tee gpurun_out/bench_elbow.json | cut -c1-200
timeout 60 python bench.py --env myoFatiLegWalk-v0 --envs-per-gpu 2048 --steps 30 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_walk.json | cut -c1-200
timeout 60 python bench.py --env myoHandObjHoldRandom-v0 --envs-per-gpu 2048 --steps 50 --warmup 3 --no-cpu-baseline | tee gpurun_out/bench_hold.json | cut -c1-200
timeout 60 python bench.py --env myoHandReachRandom-v0 --steps 100 --warmup 5 --no-cpu-baseline | tee gpurun_out/bench_reach.json | cut -c1-200
timeout 200 python bench.py --impl reference --steps 60 | tee gpurun_out/bench_ref_hand.json | cut -c1-300
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:myo_env_kernel -c 12 --csv --log-file gpurun_out/launches_myo.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 180 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 180 ncu --set full --clock-control none --import-source on -k regex:myo_env_kernel -s 4 -c 1 -f -o gpurun_out/prof_hand_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -14
