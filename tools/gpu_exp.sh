#!/bin/bash
timeout 600 python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "tests/golden")
import npg_policies
from myosuite_b200 import rollout, vec_env
pols = npg_policies.load_npz()
for env_id, n in (("myoHandObjHoldFixed-v0", 1024), ("myoHandObjHoldRandom-v0", 1024), ("myoHandReachRandom-v0", 1024), ("myoHandReachFixed-v0", 512)):
    for k in range(3):
        env = vec_env.MyoVecEnv(env_id, n, auto_reset=False, seed=11 + k)
        pol = rollout.MLPPolicy(pols[env_id][k], device=env.device)
        g = torch.Generator(device=env.device).manual_seed(k)
        _, s = rollout.examine_policy(env, pol, mode="exploration", seed=11 + k, generator=g, keep_obs=False)
        lg = pols[env_id][k]["logged"]
        print("%s policy %d: return %.1f +- %.1f (min %.1f max %.1f) success %.1f%% | logged %.1f +- %.1f (min %.1f max %.1f) success %.1f%%" % (env_id, k, s["returns"].mean(), s["returns"].std(), s["returns"].min(), s["returns"].max(), s["success_pct"], lg["stoc_pol_mean"], lg["stoc_pol_std"], lg["stoc_pol_min"], lg["stoc_pol_max"], lg["success_percentage"]))
    # untrained reference point: a zero-mean policy with the same noise
    class Z:
        def __init__(s, p): s.p = p
        def act(s, obs, mode="exploration", generator=None): return (torch.exp(s.p.log_std) * torch.randn(obs.shape[0], s.p.m, dtype=torch.float64, device=obs.device, generator=generator)).clamp(-1, 1).float()
    env = vec_env.MyoVecEnv(env_id, n, auto_reset=False, seed=5)
    _, s = rollout.examine_policy(env, Z(pol), seed=5, generator=torch.Generator(device=env.device).manual_seed(9), keep_obs=False)
    print("%s zero-mean policy: return %.1f success %.1f%%" % (env_id, s["returns"].mean(), s["success_pct"]))
PY
