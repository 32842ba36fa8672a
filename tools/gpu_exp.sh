#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_env.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra"
echo "=== default"; timeout 200 $B 2>&1 | grep "metric\|rror" | cut -c60-200
