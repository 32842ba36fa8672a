#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "dense_solver or legs or walk or contract" 2>&1 | tail -5
echo "=== legs bench"; timeout 300 python - <<'PY'
import torch, time
from myosuite_b200 import vec_env
for eid, n in (("myoFatiLegWalk-v0", 2048), ("myoLegWalk-v0", 2048)):
    env = vec_env.MyoVecEnv(eid, n, seed=0); env.reset(seed=0)
    g = torch.Generator(device=env.device).manual_seed(0)
    acts = [torch.rand(n, env.act_dim, device=env.device, generator=g)*2-1 for _ in range(8)]
    for i in range(5): env.step(acts[i % 8])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40): env.step(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/40
    print(eid, n, "envs: %.3f ms/step, %.0f env-steps/s" % (ms, n/ms*1e3))
PY
echo "=== legs phase profile"; N=2048 STEPS=20 timeout 300 python tools/gpu_phase_profile.py myoFatiLegWalk-v0 2>&1 | sed -n 2,4p | cut -c1-400
