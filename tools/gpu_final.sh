#!/bin/bash
# round-end rehearsal: GPU tests, smoke, the driver's bench command and the reference arm
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_bench.json; python - <<'PY'
import json
j = json.load(open("gpurun_out/final_bench.json"))
print({k: j[k] for k in ("value", "ms_per_step", "gpu_launches")}, j["e2e"]["value"], j["clocks"], j["roofline"]["frac"], j["cpu_baseline"]["value"])
print({k: round(v["value"]) for k, v in j["config"]["extra"].items()})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 200 python - <<'PY'
import torch
from myosuite_b200 import vec_env
for eid, n in (("myoTorsoPoseFixed-v0", 2048),):
    env = vec_env.MyoVecEnv(eid, n, seed=0); env.reset(seed=0)
    g = torch.Generator(device=env.device).manual_seed(0)
    acts = [torch.rand(n, env.act_dim, device=env.device, generator=g)*2-1 for _ in range(4)]
    for i in range(3): env.step(acts[i % 4])
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): env.step(acts[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/20
    print(eid, n, "envs: %.3f ms/step, %.0f env-steps/s" % (ms, n/ms*1e3))
PY
