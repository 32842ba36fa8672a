#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== bench torso"; timeout 200 python bench.py --env myoTorsoPoseFixed-v0 --envs-per-gpu 2048 --steps 30 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 | cut -c1-200
