#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-extra"
echo "=== default product"; timeout 100 $B 2>&1 | tail -1 | cut -c1-130
echo "=== default dbg kernel"; MYO_B200_DEBUG_KERNEL=1 timeout 100 $B 2>&1 | tail -1 | cut -c1-130
for v in oldgrad norank oldgrad_norank; do echo "=== $v"; MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_$v.so timeout 100 $B 2>&1 | tail -1 | cut -c1-130;
  echo "=== $v dbg"; MYO_B200_DEBUG_KERNEL=1 MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_$v.so timeout 100 $B 2>&1 | tail -1 | cut -c1-130; done
