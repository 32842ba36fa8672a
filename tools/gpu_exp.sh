#!/bin/bash
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra"
for k in "1,4" "1,0" "0,8" "1,12" "2,3"; do echo "=== loadkey $k"; MYO_B200_LOADKEY=$k timeout 200 $B 2>&1 | grep "metric\|rror" | cut -c60-200; done
