"""Find the first substep where GPU and oracle diverge (developer tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from myosuite_b200 import vec_env
from oracle.oracle_py import Oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from test_gpu_parity import _states, relerr
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 12; NS = int(sys.argv[2]) if len(sys.argv) > 2 else 10; NE = int(sys.argv[3]) if len(sys.argv) > 3 else 16
env = vec_env.MyoVecEnv("myoHandPoseRandom-v0", 64, taps=True, maxcon=48); m = env.mj_model; n = 64
qpos, qvel, act, ctrl = _states(m, n, np.random.default_rng(seed), overshoot=0.0, vel=0.5)
env.set_state(qpos=qpos, qvel=qvel, act=act)
ors = []
for e in range(NE):
    o = Oracle(env.I, env.D); o.reset(); o.set(qpos=qpos[e], qvel=qvel[e], act=act[e], ctrl=ctrl[e]); ors.append(o)
pmi = env.prog_info["pair_model_index"]
for s in range(NS):
    env.forward_debug(ctrl, 1); torch.cuda.synchronize()
    t = {k: v.cpu().numpy() for k, v in env.t.items() if k.startswith("tap_")}
    for e in range(NE):
        o = ors[e]; o.step(1)
        err = relerr(t["tap_qacc"][e], o.f("qacc"))
        if err > 1e-6:
            nc = int(t["tap_ncon"][e, 0])
            gp = [pmi[p] for p in t["tap_contact_pair"][e][:nc]]; gd = t["tap_contact_dist"][e][:nc]
            op = [int(p) for p in o.i("con_pair")]; od = o.f("con_dist")
            print("overflow flag", t["tap_ncon"][e, 3]); print("substep", s, "env", e, "qacc relerr %.2e" % err, "ncon gpu/oracle", nc, len(op), "niter", t["tap_ncon"][e, 2], o.solver_niter)
            for p in sorted(set(gp) | set(op)):
                g1, g2 = int(m.pair_geom1[p]), int(m.pair_geom2[p])
                a = [d for q, d in zip(gp, gd) if q == p]; b = [d for q, d in zip(op, od) if q == p]
                if not a or not b or abs(a[0] - b[0]) > 1e-9:
                    print("   pair", p, "types", m.geom_type[g1], m.geom_type[g2], "gpu dist", a, "oracle dist", b)
            sys.exit(0)
print("no divergence found in 16 envs x 10 substeps")
