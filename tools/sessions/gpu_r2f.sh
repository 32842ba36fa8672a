#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300; grep "after .* steps:\|regime:" gpurun_out/pytest_gpu.log
echo "=== bench hand"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra | tee gpurun_out/r2f_bench_hand.json | cut -c1-200
echo "=== true lone-warp phase cycles (148 envs, 1 warp per SM)"; N=148 STEPS=20 MYO_B200_WARPS_PER_CTA=1 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,4p | cut -c1-700
echo "=== 10-warp phase cycles"; STEPS=30 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,4p | cut -c1-700
