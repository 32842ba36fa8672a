"""Trained-policy replay on the CPU oracle (TEST INFRASTRUCTURE): the strongest anchor to real MuJoCo output the reference holds.

The reference commits its NPG baselines -- policies trained for 1000 iterations on MuJoCo + the reference env code -- together with
their logged returns (agents/baslines_NPG/<env>/*/*/{iterations/best_policy.pickle, logs/log.csv}).  A trained policy is tuned to
the dynamics it saw: replaying it on this repo's physics and recovering the logged return is evidence that the physics restatement
is faithful where the policy operates (a random policy's return is dominated by the reward's distance term and says much less).
Protocol restated from mjrl (sample_paths / GymEnv.step): horizon = max_episode_steps, action = mean + exp(log_std) N(0,1)
(evaluation: the mean), clipped to the action space, episode ends on done; success = sum(solved) > 5 (env_base.evaluate_success).

    python tests/devtools/policy_replay.py ENV_ID [n_traj] [stochastic|mean] [run index]
"""
import copy
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def rollouts(args):
    env_id, n, seed, pol, mode, clip = args
    import npg_policies
    from myosuite_b200 import assets, blob, mjcf, vec_env
    from oracle import env_oracle
    from oracle.oracle_py import Oracle
    T, kw, entry = vec_env.env_spec(env_id)
    m = assets.load(vec_env._MODEL_OF_XML[kw["model_path"]])
    o = Oracle(*blob.pack(m))
    rng = np.random.default_rng(seed)
    dt = m.opt_timestep * 10
    task = "pose" if "pose_v0" in entry else "reach" if "reach_v0" in entry else "hold"
    sigma = np.exp(pol["log_std"])
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
        if task == "hold" and "Random" in env_id:        # ObjHoldRandomEnvV0.reset (obj_hold_v0.py:126-145)
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
        tm = 0.0

        def observe():
            if task == "pose":
                return env_oracle.pose_obs(o.f("qpos"), o.f("qvel"), o.f("act"), tgt, dt), env_oracle.pose_reward(o.f("qpos").copy(), o.f("act").copy(), tgt, thd)
            o.forward()
            if task == "reach":
                return env_oracle.reach_obs_reward(m, o, dt, tips, tg, tm, far_th=kw.get("far_th", 0.35))
            ob_, r_ = env_oracle.hold_obs_reward(m, o, dt, goal); r_["solved"] = bool(-r_["goal_dist"] < 0.010)
            return ob_, r_
        obs, _ = observe()
        R = solved = 0
        for t in range(T):
            a = npg_policies.mean_action(pol, obs.astype(np.float64))
            if mode == "stochastic":
                a = a + sigma * rng.normal(0, 1, m.nu)
            if clip:
                a = np.clip(a, -1.0, 1.0)
            env_oracle.env_step(o, a, 10)
            for _s in range(10):
                tm += m.opt_timestep
            obs, r = observe()
            R += r["dense"]; solved += float(r["solved"])
            if r["done"]:
                break
        out.append((R, solved))
    return out


def run(env_id, n_traj, mode="stochastic", run_index=0, clip=True, procs=None, seed=0):
    import npg_policies
    pol = npg_policies.load_npz()[env_id][run_index]
    procs = procs or min(os.cpu_count() or 1, 8)
    per = [n_traj // procs + (1 if i < n_traj % procs else 0) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        res = sum(pool.map(rollouts, [(env_id, per[i], seed * 1000 + i, pol, mode, clip) for i in range(procs) if per[i]]), [])
    R = np.array([r[0] for r in res]); S = np.array([r[1] for r in res])
    return dict(n=len(R), mean=float(R.mean()), std=float(R.std()), max=float(R.max()), min=float(R.min()), success_pct=float(100 * np.mean(S > 5)),
                logged={k: round(float(v), 2) for k, v in pol["logged"].items()})


if __name__ == "__main__":
    env_id = sys.argv[1] if len(sys.argv) > 1 else "myoElbowPose1D6MFixed-v0"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    mode = sys.argv[3] if len(sys.argv) > 3 else "stochastic"
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    for clip in (True, False):
        print(env_id, mode, "run", k, "clip", clip, json.dumps(run(env_id, n, mode, k, clip)))
