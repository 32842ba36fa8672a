#!/bin/bash
# round-2 session A: correctness of the shared-space / row-register-Cholesky kernel + throughput + knob sweep + phase cycles
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | cut -c1-400
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench hand"; timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2a_bench_hand.json | cut -c1-330
for w in 8 9; do echo "=== warps $w"; MYO_B200_WARPS_PER_CTA=$w timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline | cut -c1-160; done
echo "=== solve_sync 0"; MYO_B200_SOLVE_SYNC=0 timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline | cut -c1-160
echo "=== dbg kernel"; MYO_B200_DEBUG_KERNEL=1 timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline | cut -c1-160
echo "=== lockstep groups 2 (dbg kernel)"; timeout 100 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --lockstep-groups 2 | cut -c1-160
echo "=== other configs"
timeout 60 python bench.py --env myoElbowPose1D6MRandom-v0 --steps 300 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2a_bench_elbow.json | cut -c1-160
timeout 80 python bench.py --env myoFatiLegWalk-v0 --envs-per-gpu 2048 --steps 30 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2a_bench_walk.json | cut -c1-160
timeout 80 python bench.py --env myoHandObjHoldRandom-v0 --envs-per-gpu 2048 --steps 50 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2a_bench_hold.json | cut -c1-160
echo "=== phase cycles"; timeout 200 python tools/gpu_phase_profile.py 2>&1 | tail -9 | cut -c1-600
echo "=== phase waits"; WAITS=1 timeout 200 python tools/gpu_phase_profile.py 2>&1 | sed -n 2,3p | cut -c1-600
