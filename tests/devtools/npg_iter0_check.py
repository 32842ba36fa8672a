"""Statistical drop-in check against numbers the REFERENCE STACK produced: iteration 0 of the reference's
NPG baselines (myosuite/agents/baslines_NPG/<env>/*/*/logs/log.csv) sampled 96 trajectories of a freshly initialised Gaussian MLP
policy (output layer scaled 1e-2 => mean action ~ 0, init_log_std = -0.25) on MuJoCo + the reference env code, and logged the mean /
std / max / min return and the success rate.  This script replays that protocol on this repo's CPU oracle + env logic (test
infrastructure, hence under tests/; no GPU) and prints the same statistics:   python tests/devtools/npg_iter0_check.py [env_id] [n_traj]
Reference values (3 seeds each, nearly identical because the env seeds coincide):
  myoElbowPose1D6MRandom-v0: mean 65.5 / 61.3 / 63.0, std 191 / 189 / 191, max 632.9, min -162.3, success 55.2 / 54.2 / 54.2 %
  myoHandPoseRandom-v0:      mean -336.1, std 28.0-28.2, max -265.8, min -408.6, success 0 %
  myoHandPoseFixed-v0:       mean -407.5, std 10.3-10.7, max -377.6, min -433.9
  myoElbowPose1D6MFixed-v0:  mean 40.4-43.3, std 84.1-84.8, max 236.0, min -198.2, success 69.8-71.9 %
  myoHandReachRandom-v0:     mean -30.0, std 7.6-7.7, max -7.27, min -42.84     myoHandReachFixed-v0: mean -20.6, std 7.6-8.0, max 14.73, min -35.89
  myoHandObjHoldRandom-v0:   mean -244.7, std 28.2-28.9, max -184.8, min -315..-330   myoHandObjHoldFixed-v0: mean -224.6, std 15.2-15.9, max -181, min -265.8
(the baselines date from Feb 2022, myosuite v0.2dev: env definitions that changed since then cannot match)
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def rollouts(args):
    env_id, n, seed = args
    import copy
    from myosuite_b200 import assets, blob, vec_env, mjcf
    from oracle import env_oracle
    from oracle.oracle_py import Oracle
    T, kw, entry = vec_env.env_spec(env_id)
    m = assets.load(vec_env._MODEL_OF_XML[kw["model_path"]])
    o = Oracle(*blob.pack(m))
    rng = np.random.default_rng(seed)
    sigma, dt = np.exp(-0.25), m.opt_timestep * 10
    task = "pose" if "pose_v0" in entry else "reach" if "reach_v0" in entry else "hold"
    if task == "pose":
        thd = kw["pose_thd"]; lo, hi = np.zeros(m.nq), np.zeros(m.nq)
        if kw.get("target_jnt_range"):
            for jn, (a, b) in kw["target_jnt_range"].items():
                qa = m.jnt_qposadr[m.name2id("joint", jn)]; lo[qa], hi[qa] = a, b
        else:
            lo = hi = np.asarray(kw["target_jnt_value"], dtype=np.float64)
    if task == "reach":
        tips = list(kw["target_reach_range"].keys())
    if task == "hold":
        kin = mjcf.kinematics(m, m.qpos0); ob, sid = m.name2id("body", "object"), m.name2id("site", "object")
        obj_init = kin["xpos"][ob] + kin["xmat"][ob] @ m.site_pos[sid]
        q0 = m.qpos0.copy(); q0[:-7] = 0.0; q0[0] = -1.5                                  # obj_hold_v0.py:61-62
        gid = m.name2id("geom", "object")
    out = []
    for _ in range(n):
        if task == "hold" and "Random" in env_id:        # ObjHoldRandomEnvV0.reset (obj_hold_v0.py:126-145): new goal, new object size
            m2 = copy.deepcopy(m); m2.geom_size[gid] = rng.uniform(0.020, 0.030, 3)
            o = Oracle(*blob.pack(m2)); goal = obj_init + rng.uniform(-0.030, 0.030, 3)
        elif task == "hold":
            goal = m.site_pos[m.name2id("site", "goal")].copy()
        o.reset()
        if task == "pose":
            tgt = rng.uniform(lo, hi)
            q = np.array([rng.uniform(*m.jnt_range[j]) for j in range(m.njnt)]) if kw.get("reset_type", "init") == "random" else m.qpos0.copy()
            o.set(qpos=q, qvel=np.zeros(m.nv), act=np.zeros(m.na))
        elif task == "reach":
            tg = np.array([rng.uniform(kw["target_reach_range"][t][0], kw["target_reach_range"][t][1]) for t in tips])
            o.set(qpos=m.qpos0, qvel=np.zeros(m.nv), act=np.zeros(m.na))
        else:
            o.set(qpos=q0, qvel=np.zeros(m.nv), act=np.zeros(m.na))
        R = solved = 0; tm = 0.0
        for t in range(T):
            env_oracle.env_step(o, rng.normal(0, sigma, m.nu), 10)
            for _s in range(10):
                tm += m.opt_timestep
            if task == "pose":
                r = env_oracle.pose_reward(o.f("qpos").copy(), o.f("act").copy(), tgt, thd)
            elif task == "reach":
                o.forward(); _, r = env_oracle.reach_obs_reward(m, o, dt, tips, tg, tm, far_th=kw.get("far_th", 0.35)); r["solved"] = float(r["solved"])
            else:
                o.forward(); _, r = env_oracle.hold_obs_reward(m, o, dt, goal); r["solved"] = 0.0
            R += r["dense"]; solved += r["solved"]
            if r["done"]:
                break
        out.append((R, solved))
    return out


def run(env_id, n_traj, procs=None, seed=0):
    procs = procs or min(os.cpu_count() or 1, 8)
    per = [n_traj // procs + (1 if i < n_traj % procs else 0) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        res = sum(pool.map(rollouts, [(env_id, per[i], seed * 1000 + i) for i in range(procs) if per[i]]), [])
    R = np.array([r[0] for r in res]); S = np.array([r[1] for r in res])
    return dict(n=len(R), mean=float(R.mean()), std=float(R.std()), max=float(R.max()), min=float(R.min()), success_pct=float(100 * np.mean(S > 5)))


if __name__ == "__main__":
    env_id = sys.argv[1] if len(sys.argv) > 1 else "myoElbowPose1D6MRandom-v0"
    print(env_id, json.dumps(run(env_id, int(sys.argv[2]) if len(sys.argv) > 2 else 96)))
