#!/bin/bash
mkdir -p gpurun_out
echo "=== bench default (driver command)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 | tee gpurun_out/r2h_bench_default.json | cut -c1-2500
echo "=== reference arm"; timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 | tee gpurun_out/r2h_bench_reference.json | cut -c1-1200
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
