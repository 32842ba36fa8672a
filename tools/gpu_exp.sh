#!/bin/bash
B="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra"
for v in 41 40 01 00; do echo "=== barrier mask 0x$v"; MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_bm$v.so timeout 200 $B 2>&1 | grep "metric\|rror" | cut -c60-200; done
echo "=== 0x41 on hold/legs"; MYO_B200_LIB=$PWD/myosuite_b200/libmyo_b200_bm41.so timeout 300 python - <<'PY'
import torch
from myosuite_b200 import vec_env
for eid, n in (("myoHandObjHoldRandom-v0", 2048), ("myoFatiLegWalk-v0", 2048), ("myoElbowPose1D6MRandom-v0", 4096)):
    env = vec_env.MyoVecEnv(eid, n, seed=0); env.reset(seed=0)
    g = torch.Generator(device=env.device).manual_seed(0)
    acts = [torch.rand(n, env.act_dim, device=env.device, generator=g)*2-1 for _ in range(4)]
    for i in range(3): env.step(acts[i % 4])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): env.step(acts[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/30
    print(eid, n, "envs: %.3f ms/step, %.0f env-steps/s" % (ms, n/ms*1e3))
PY
