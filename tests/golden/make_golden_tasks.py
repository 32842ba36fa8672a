"""Golden vectors for the Walk / ObjHold task logic, produced by RUNNING THE REFERENCE'S OWN CLASSES
(WalkEnvV0 / ObjHoldFixedEnvV0 get_obs_dict, get_reward_dict and helpers, unmodified, created with __new__ so that no
simulator is constructed) on mj_data-like records filled from this repo's CPU oracle.  Output: tests/golden/tasks.npz
Run here (needs /root/reference and oracle/libmyo_oracle.so):  python tests/golden/make_golden_tasks.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import _ref_stubs  # noqa: E402

_ref_stubs.install()
from myosuite.envs.myo.myobase.obj_hold_v0 import ObjHoldFixedEnvV0  # noqa: E402
from myosuite.envs.myo.myobase.walk_v0 import WalkEnvV0  # noqa: E402
from myosuite.envs.obs_vec_dict import ObsVecDict  # noqa: E402

from myosuite_b200 import assets, blob  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

out = {}
rng = np.random.default_rng(77)


def fake_model(m):
    fm = types.SimpleNamespace(na=m.na, nu=m.nu, body_mass=m.body_mass, jnt_qposadr=m.jnt_qposadr, key_qpos=m.key_qpos,
                               opt=types.SimpleNamespace(timestep=m.opt_timestep))
    fm.body = lambda name: types.SimpleNamespace(id=m.name2id("body", name))
    fm.joint = lambda name: types.SimpleNamespace(id=m.name2id("joint", name))
    fm.site = lambda name: types.SimpleNamespace(id=m.name2id("site", name))
    return fm


def fake_data(o, nb):
    return types.SimpleNamespace(time=0.0, qpos=o.f("qpos").copy(), qvel=o.f("qvel").copy(), act=o.f("act").copy(), xpos=o.f("xpos").reshape(nb, 3).copy(),
                                 xquat=o.f("xquat").reshape(nb, 4).copy(), xipos=o.f("xipos").reshape(nb, 3).copy(), cvel=o.f("cvel").reshape(nb, 6).copy(),
                                 actuator_length=o.f("actuator_length").copy(), actuator_velocity=o.f("actuator_velocity").copy(),
                                 actuator_force=o.f("actuator_force").copy(), site_xpos=o.f("site_xpos").reshape(-1, 3).copy())


# ---- Walk: states = keyframes + noise, a few oracle steps in between to get velocities / muscle states
m = assets.load("myolegs")
o = Oracle(*blob.pack(m))
N = 24
S = dict(qpos=[], qvel=[], act=[], steps=[], obs=[], dense=[], done=[], vel_reward=[], cyclic_hip=[], ref_rot=[], joint_angle_rew=[])
for i in range(N):
    o.reset()
    q = m.key_qpos[i % 4].copy(); q[7:] += rng.normal(0, 0.05, m.nq - 7)
    if i >= 16:                                      # tilted / low states so that the done branches fire
        q[2] = 0.7 if i % 2 else 0.95
        ang = rng.uniform(0.5, 1.4); q[3:7] = [np.cos(ang / 2), 0, np.sin(ang / 2) * (i % 2), np.sin(ang / 2) * (1 - i % 2)]
    o.set(qpos=q, qvel=rng.normal(0, 0.5, m.nv), act=rng.uniform(0, 1, m.na), ctrl=np.zeros(m.nu))
    o.forward()                                      # the reference's observed data: forward on the state with ctrl = 0 (robot.py:595-607)
    steps = int(rng.integers(0, 300))
    env = WalkEnvV0.__new__(WalkEnvV0)
    env.mj_model, env.mj_data = fake_model(m), fake_data(o, m.nbody)
    env.frame_skip, env.steps, env.hip_period, env.min_height, env.max_rot = 10, steps, 100, 0.8, 0.8      # dt = timestep * frame_skip = 0.01
    env.target_x_vel, env.target_y_vel, env.target_rot, env.init_qpos = 0.0, 1.2, None, m.key_qpos[0].copy()
    env.rwd_keys_wt = WalkEnvV0.DEFAULT_RWD_KEYS_AND_WEIGHTS
    od = env.get_obs_dict(env.mj_model, env.mj_data)
    env.obs_dict = od
    ovd = ObsVecDict()
    t, vec = ovd.obsdict2obsvec(od, WalkEnvV0.DEFAULT_OBS_KEYS + ["act"])      # base_v0.py:33-37 appends "act"
    rd = env.get_reward_dict(od)
    S["qpos"].append(o.f("qpos").copy()); S["qvel"].append(o.f("qvel").copy()); S["act"].append(o.f("act").copy()); S["steps"].append(steps); S["obs"].append(vec)
    for k in ("dense", "done", "vel_reward", "cyclic_hip", "ref_rot", "joint_angle_rew"):
        S[k].append(float(np.asarray(rd[k]).ravel()[0]))
for k, v in S.items():
    out["walk_" + k] = np.array(v)
assert out["walk_obs"].shape == (N, 403) and out["walk_obs"].dtype == np.float32 and out["walk_done"].sum() >= 2

# ---- ObjHold
m = assets.load("myohand_hold")
o = Oracle(*blob.pack(m))
N = 16
H = dict(qpos=[], qvel=[], act=[], goal=[], obs=[], dense=[], done=[])
for i in range(N):
    o.reset()
    q = m.qpos0.copy(); q[:-7] = rng.uniform(m.jnt_range[:23, 0], m.jnt_range[:23, 1]); q[-7:-4] += rng.normal(0, 0.01 if i < 12 else 0.4, 3)
    o.set(qpos=q, qvel=rng.normal(0, 0.5, m.nv), act=rng.uniform(0, 1, m.na)); o.forward()
    goal = np.array([-.240, -.520, 1.470]) + rng.uniform(-0.03, 0.03, 3)
    env = ObjHoldFixedEnvV0.__new__(ObjHoldFixedEnvV0)
    env.mj_model, env.mj_data, env.frame_skip = fake_model(m), fake_data(o, m.nbody), 10                 # dt = 0.02
    env.object_sid, env.goal_sid = m.name2id("site", "object"), m.name2id("site", "goal")
    env.mj_data.site_xpos[env.goal_sid] = goal
    env.rwd_keys_wt = ObjHoldFixedEnvV0.DEFAULT_RWD_KEYS_AND_WEIGHTS
    od = env.get_obs_dict(env.mj_model, env.mj_data)
    env.obs_dict = od
    t, vec = ObsVecDict().obsdict2obsvec(od, ObjHoldFixedEnvV0.DEFAULT_OBS_KEYS + ["act"])
    rd = env.get_reward_dict(od)
    H["qpos"].append(o.f("qpos").copy()); H["qvel"].append(o.f("qvel").copy()); H["act"].append(o.f("act").copy()); H["goal"].append(goal); H["obs"].append(vec)
    H["dense"].append(float(np.asarray(rd["dense"]).ravel()[0])); H["done"].append(float(np.asarray(rd["done"]).ravel()[0]))
for k, v in H.items():
    out["hold_" + k] = np.array(v)
assert out["hold_obs"].shape == (N, 91) and out["hold_done"].sum() >= 1
np.savez_compressed(os.path.join(HERE, "tasks.npz"), **out)
print("wrote tasks.npz", {k: v.shape for k, v in out.items()})
