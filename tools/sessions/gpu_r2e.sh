#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extra"
echo "=== default"; timeout 100 $B 2>&1 | tail -1 | cut -c1-130
for v in nomerge nobatch nomerge_nobatch; do echo "=== $v"; MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_$v.so timeout 100 $B 2>&1 | tail -1 | cut -c1-130; done
echo "=== ncu full (1 launch)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:myo_env_kernel -s 6 -c 1 -f -o gpurun_out/r02_hand_mid python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
