#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command's timed region + one full capture of the step kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r02_launches_timed_region.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:myo_env_kernel -s 6 -c 1 -f -o gpurun_out/r02_hand_final python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
echo "=== bench hand (same build, not under ncu)"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra | tee gpurun_out/r2x_bench_hand.json | cut -c1-200
