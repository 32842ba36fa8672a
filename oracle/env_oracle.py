"""CPU ORACLE for the Python-side part of the hot path (TEST INFRASTRUCTURE, not product code).

numpy restatement of what the reference does around mj_step each control step; every function cites
the reference lines it follows.  Pinned against tests/golden/pylogic.npz, which was produced by running
the reference's own unmodified code (tests/golden/make_golden.py).
"""
import numpy as np


def action_to_ctrl(a, normalize_act=True):
    """/root/reference/myosuite/envs/myo/base_v0.py:83-96 (all actuators of the hot-path models are muscles)."""
    a = np.array(a, dtype=np.float64, copy=True)
    if normalize_act:
        a = 1.0 / (1.0 + np.exp(-5.0 * (a - 0.5)))
    return a


class Fatigue:
    """3CC-r cumulative fatigue, /root/reference/myosuite/envs/myo/fatigue.py:8-99."""

    def __init__(self, n, dt, tauact=0.01, taudeact=0.04):
        self.r, self.F, self.R = 10 * 15, 0.00912, 0.1 * 0.00094      # fatigue.py:9-11
        self.dt = dt                                                 # fatigue.py:12 (timestep * frame_skip)
        self.tauact, self.taudeact = np.broadcast_to(tauact, (n,)).astype(float), np.broadcast_to(taudeact, (n,)).astype(float)
        self.MA, self.MR, self.MF = np.zeros(n), np.ones(n), np.zeros(n)   # fatigue.py:18-20

    def compute_act(self, act):
        """fatigue.py:38-76 (order of the masked assignments matters; returns the updated MA as the muscle ctrl)."""
        TL = np.array(act, dtype=np.float64)
        MA, MR, MF = self.MA, self.MR, self.MF
        LD = 1 / self.tauact * (0.5 + 1.5 * MA)
        LR = (0.5 + 1.5 * MA) / self.taudeact
        C = np.zeros_like(MA)
        i = (MA < TL) & (MR > (TL - MA)); C[i] = LD[i] * (TL[i] - MA[i])
        i = (MA < TL) & (MR <= (TL - MA)); C[i] = LD[i] * MR[i]
        i = MA >= TL; C[i] = LR[i] * (TL[i] - MA[i])
        rR = np.where(MA >= TL, self.r * self.R, self.R)
        C = np.clip(C, np.maximum(-MA / self.dt + self.F * MA, (MR - 1) / self.dt + rR * MF),
                    np.minimum((1 - MA) / self.dt + self.F * MA, MR / self.dt + rR * MF))
        dMA, dMR, dMF = (C - self.F * MA) * self.dt, (-C + rR * MF) * self.dt, (self.F * MA - rR * MF) * self.dt
        self.MA, self.MR, self.MF = MA + dMA, MR + dMR, MF + dMF
        return self.MA, self.MR, self.MF


def pose_obs(qpos, qvel, act, target, dt):
    """/root/reference/myosuite/envs/myo/myobase/pose_v0.py:100-111 + envs/obs_vec_dict.py:76-88
    (keys qpos, qvel, pose_err, act in that order; float32)."""
    return np.concatenate([qpos, np.asarray(qvel) * dt, np.asarray(target) - qpos, act]).astype(np.float32)


def pose_reward(qpos, act, target, pose_thd, weights=(1.0, 4.0, 1.0, 50.0)):
    """/root/reference/myosuite/envs/myo/myobase/pose_v0.py:113-140 -> dict(dense, solved, done, pose, bonus, penalty, act_reg)."""
    pose_dist = np.linalg.norm(np.asarray(target) - qpos)
    act_mag = np.linalg.norm(act) / (len(act) if len(act) else 1)
    far_th = 4 * np.pi / 2
    r = dict(pose=-pose_dist, bonus=1.0 * (pose_dist < pose_thd) + 1.0 * (pose_dist < 1.5 * pose_thd), penalty=-1.0 * (pose_dist > far_th),
             act_reg=-act_mag, sparse=-pose_dist, solved=pose_dist < pose_thd, done=pose_dist > far_th)
    r["dense"] = weights[0] * r["pose"] + weights[1] * r["bonus"] + weights[2] * r["act_reg"] + weights[3] * r["penalty"]
    return r


def env_step(oracle, action, frame_skip, normalize_act=True, fatigue=None):
    """One control step on the physics oracle: base_v0.py:82-118 + robot.py:901-905 (same ctrl for all substeps)."""
    ctrl = action_to_ctrl(action, normalize_act)
    if fatigue is not None:
        ctrl = fatigue.compute_act(ctrl)[0].copy()
    oracle.set(ctrl=ctrl)
    oracle.step(frame_skip)
    return ctrl
