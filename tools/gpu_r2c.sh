#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | cut -c1-300
echo "=== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-300
echo "=== bench hand"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra | tee gpurun_out/r2c_bench_hand.json | cut -c1-200
echo "=== lone-warp phase cycles"; N=1184 STEPS=20 MYO_B200_WARPS_PER_CTA=1 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,4p | cut -c1-700
echo "=== 10-warp phase cycles"; STEPS=30 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,4p | cut -c1-700
