#!/bin/bash
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extra --maxcon 16"
for w in 8 10 12 13; do echo "=== r128 maxcon16 warps $w"; MYO_B200_PRODUCT_ONLY=1 MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_r128.so MYO_B200_WARPS_PER_CTA=$w timeout 100 $B 2>&1 | tail -1 | cut -c1-130; done
for w in 8 10; do echo "=== default(r168) maxcon16 warps $w"; MYO_B200_WARPS_PER_CTA=$w timeout 100 $B 2>&1 | tail -1 | cut -c1-130; done
