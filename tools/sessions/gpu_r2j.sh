#!/bin/bash
mkdir -p gpurun_out
echo "=== chol ubench"; timeout 60 tools/ubench/ubench_chol
echo "=== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench hand"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra | tee gpurun_out/r2j_bench_hand.json | cut -c1-200
echo "=== 10-warp phase cycles"; STEPS=30 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,4p | cut -c1-700
