#!/bin/bash
mkdir -p gpurun_out
echo "=== chol ubench"; timeout 60 tools/ubench/ubench_chol
echo "=== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -14 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench hand"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra | tee gpurun_out/r2c_bench_hand.json | cut -c1-200
