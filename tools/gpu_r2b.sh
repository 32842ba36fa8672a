#!/bin/bash
# round-2 session B: register-budget x warps-per-CTA matrix, correctness, lone-warp phase profile
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline"
for v in r168 r200 r144; do for w in 7 8 9 10; do
  echo "=== $v warps $w"; MYO_B200_PRODUCT_ONLY=1 MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_$v.so MYO_B200_WARPS_PER_CTA=$w timeout 100 $B 2>&1 | tail -1 | cut -c1-130
done; done
echo "=== default lib, default warps"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2b_bench_hand.json | cut -c1-200
echo "=== other configs"
timeout 60 python bench.py --env myoElbowPose1D6MRandom-v0 --steps 300 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2b_bench_elbow.json | cut -c1-160
timeout 80 python bench.py --env myoFatiLegWalk-v0 --envs-per-gpu 2048 --steps 30 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2b_bench_walk.json | cut -c1-160
timeout 80 python bench.py --env myoHandObjHoldRandom-v0 --envs-per-gpu 2048 --steps 50 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2b_bench_hold.json | cut -c1-160
echo "=== lone-warp phase cycles (1 warp per CTA, no lockstep partners)"; N=1184 STEPS=20 MYO_B200_WARPS_PER_CTA=1 timeout 200 python tools/gpu_phase_profile.py 2>&1 | tail -9 | cut -c1-700
echo "=== 10-warp phase cycles"; STEPS=30 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,4p | cut -c1-700
echo "=== 10-warp phase waits"; WAITS=1 STEPS=30 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,3p | cut -c1-700
