#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | grep "worst\|passed\|failed\|FAILED\|Error\|assert [0-9]" | tail -30
echo "=== driver cmd"; MYO_B200_VERBOSE=1 timeout 400 python bench.py --steps 20 --warmup 3 2> gpurun_out/r2n_bench.err | tail -1 > gpurun_out/r2n_bench.json; grep "myo_b200" gpurun_out/r2n_bench.err | cut -c1-200; grep -o '"value": [0-9.]*\|"clocks": {[^}]*}\|"e2e": {[^}]*}\|"myo[A-Za-z0-9-]*": {"envs_per_gpu": [0-9]*, "steps": [0-9]*, "value": [0-9.]*' gpurun_out/r2n_bench.json
echo "=== 100 steps"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extra | tee gpurun_out/r2n_bench100.json | cut -c1-200
