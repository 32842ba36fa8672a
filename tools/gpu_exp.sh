#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "Elbow or elbow or env or contract" 2>&1 | tail -3
timeout 300 python - <<'PY'
import torch, os
from myosuite_b200 import vec_env
os.environ["MYO_B200_VERBOSE"]="1"
for eid, n in (("myoElbowPose1D6MRandom-v0", 4096), ("myoElbowPose1D6MRandom-v0", 16384), ("myoElbowPose1D6MRandom-v0", 65536)):
    env = vec_env.MyoVecEnv(eid, n, seed=0); env.reset(seed=0)
    g = torch.Generator(device=env.device).manual_seed(0)
    acts = [torch.rand(n, env.act_dim, device=env.device, generator=g)*2-1 for _ in range(8)]
    for i in range(5): env.step(acts[i % 8])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): env.step(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/100
    print(eid, n, "envs: %.3f ms/step, %.0f env-steps/s" % (ms, n/ms*1e3))
PY
