#!/bin/bash
mkdir -p gpurun_out
echo "=== icache ubench"; timeout 120 tools/ubench/ubench_icache | tee gpurun_out/r2d_ubench_icache.txt
echo "=== parity tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gym_api.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300; grep "after .* steps:\|regime:" gpurun_out/pytest_gpu.log
