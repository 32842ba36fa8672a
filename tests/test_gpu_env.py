"""GPU tests of the batched env behaviour at BASELINE sizes: size-independent properties (determinism, batch-size
independence, auto-reset, TimeLimit, bounded observations) -- the oracle is too slow for 4096 envs x many steps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(eid, n, steps, seed, act_seed=0, n_act=4096, **kw):
    import torch
    from myosuite_b200 import vec_env
    env = vec_env.MyoVecEnv(eid, n, seed=seed, **kw)
    env.reset(seed=seed)
    g = torch.Generator(device="cpu").manual_seed(act_seed)
    out = []
    for _ in range(steps):
        # env i always receives row i of the same [n_act, nu] random matrix, whatever the batch size
        a = (torch.rand(n_act, env.act_dim, generator=g) * 2 - 1)[:n].contiguous().to(env.device)
        obs, rew, done, trunc, info = env.step(a)
        out.append((obs.clone(), rew.clone(), done.clone(), trunc.clone()))
    torch.cuda.synchronize()
    return env, out


@pytest.mark.parametrize("eid", ["myoElbowPose1D6MRandom-v0", "myoHandPoseRandom-v0"])
def test_full_size_rollout_properties(eid):
    import torch
    n, steps = 4096, 110
    env, out = _run(eid, n, steps, seed=3)
    obs = torch.stack([o[0] for o in out]); rew = torch.stack([o[1] for o in out])
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    m = env.mj_model
    act = obs[..., 2 * m.nq + m.nv:]
    assert act.min() >= -1e-6 and act.max() <= 1 + 1e-6                       # activations stay in [0,1]
    # TimeLimit: with no early termination every env truncates exactly at step 100 and is auto-reset
    trunc = torch.stack([o[3] for o in out]); done = torch.stack([o[2] for o in out])
    first = (trunc | done).float().argmax(0)
    assert ((trunc | done).sum(0) >= 1).all()
    assert (first[~done.bool().any(0)] == env.max_episode_steps - 1).all()
    assert (env.t["episode_count"] >= 2).all()
    # after auto-reset the obs is the reset obs: qvel*dt == 0 and act == 0 for envs reset on the last step before
    k = env.max_episode_steps - 1
    o_reset = out[k][0]
    assert torch.all(o_reset[:, m.nq:m.nq + m.nv][trunc[k].bool()] == 0)
    assert torch.all(o_reset[:, 2 * m.nq + m.nv:][trunc[k].bool()] == 0)
    # reset poses are inside the joint ranges; targets inside the target ranges
    q0 = o_reset[:, :m.nq][trunc[k].bool()].double().cpu().numpy()
    lo, hi = m.jnt_range[:, 0], m.jnt_range[:, 1]
    assert np.all(q0 >= lo - 1e-6) and np.all(q0 <= hi + 1e-6)
    tr = env.t["target_range"].cpu().numpy(); tg = env.t["target"].cpu().numpy()
    assert np.all(tg >= tr[:, 0] - 1e-12) and np.all(tg <= tr[:, 1] + 1e-12)
    assert env.t["last_return"].abs().sum() > 0


def test_determinism_and_batch_size_independence():
    """Same seed -> bit-identical rollouts; env i's trajectory does not depend on how many envs share the launch."""
    import torch
    eid = "myoHandPoseRandom-v0"
    _, a = _run(eid, 256, 12, seed=5)
    _, b = _run(eid, 256, 12, seed=5)
    for (oa, ra, da, ta), (ob, rb, db, tb) in zip(a, b):
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    env_s, c = _run(eid, 64, 12, seed=5)
    for (oa, ra, _, _), (oc, rc, _, _) in zip(a, c):
        assert torch.equal(oa[:64], oc) and torch.equal(ra[:64], rc)
    _, d = _run(eid, 256, 12, seed=6)
    assert not torch.equal(a[0][0], d[0][0])


def test_load_sorted_rounds_do_not_change_results(monkeypatch):
    """The env -> (CTA, round) map of the product kernel is re-sorted by contact load before every step (myo_regroup_kernel); each env's
    trajectory must be bit-identical to the fixed e % grid map, for a batch that fills two rounds and for one with idle slots."""
    import torch
    eid = "myoHandPoseRandom-v0"
    for n in (4096, 300):
        monkeypatch.setenv("MYO_B200_REGROUP", "1"); env_a, a = _run(eid, n, 8, seed=11)
        monkeypatch.setenv("MYO_B200_REGROUP", "0"); env_b, b = _run(eid, n, 8, seed=11)
        for (oa, ra, da, ta), (ob, rb, db, tb) in zip(a, b):
            assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
        assert torch.equal(env_a.t["qpos"], env_b.t["qpos"]) and torch.equal(env_a.t["qvel"], env_b.t["qvel"])
        assert env_a.batch.launches == env_b.batch.launches + 8               # one regroup launch per env step


def test_env_offset_shards_reproduce_global_batch():
    """Multi-GPU sharding rule: rank r simulates envs [r*n, (r+1)*n) with env_offset = r*n and gets exactly the
    rows a single big batch would have produced (no data-path collective needed)."""
    import torch
    from myosuite_b200 import vec_env
    eid, n = "myoElbowPose1D6MRandom-v0", 128
    big = vec_env.MyoVecEnv(eid, 2 * n, seed=9); big.reset(seed=9)
    lo = vec_env.MyoVecEnv(eid, n, seed=9, env_offset=0); lo.reset(seed=9)
    hi = vec_env.MyoVecEnv(eid, n, seed=9, env_offset=n); hi.reset(seed=9)
    g = torch.Generator().manual_seed(1)
    for _ in range(20):
        a = (torch.rand(2 * n, big.act_dim, generator=g) * 2 - 1).cuda()
        ob, rb, *_ = big.step(a); ol, rl, *_ = lo.step(a[:n].contiguous()); oh, rh, *_ = hi.step(a[n:].contiguous())
        assert torch.equal(ob[:n], ol) and torch.equal(ob[n:], oh) and torch.equal(rb[n:], rh)


def test_single_env_gym_facade():
    """make(id) -> reset()/step() with the reference's call shapes and dtypes (env_base.py:403-407,647-654)."""
    from myosuite_b200 import make
    env = make("myoElbowPose1D6MRandom-v0", seed=1234)
    obs, info = env.reset(seed=1234)
    assert obs.dtype == np.float32 and obs.shape == (9,) and isinstance(info, dict)
    o2, r, term, trunc, info = env.step(np.zeros(6, dtype=np.float32))
    assert o2.shape == (9,) and isinstance(r, float) and isinstance(term, bool) and trunc is False
    assert env.unwrapped.mj_model.nu == 6 and env.unwrapped.mj_model.na == 6 and env.dt == pytest.approx(0.02)
    # determinism check of the reference's test strategy (tests/test_envs.py:100-121): same seed, same reset obs / step obs
    env2 = make("myoElbowPose1D6MRandom-v0", seed=1234)
    obs_b, _ = env2.reset(seed=1234)
    np.testing.assert_allclose(obs, obs_b, atol=1e-5)
    o2b, rb, *_ = env2.step(np.zeros(6, dtype=np.float32))
    np.testing.assert_allclose(o2, o2b, atol=1e-5); assert r == pytest.approx(rb, abs=1e-5)


def test_sarcopenia_variant_halves_gain():
    import torch
    from myosuite_b200 import vec_env
    a = vec_env.MyoVecEnv("myoElbowPose1D6MRandom-v0", 4, taps=True)
    b = vec_env.MyoVecEnv("myoSarcElbowPose1D6MRandom-v0", 4, taps=True)
    q = np.full((4, 1), 1.0); act = np.ones((4, 6)); ctrl = np.ones((4, 6))
    for e in (a, b):
        e.set_state(qpos=q, qvel=np.zeros((4, 1)), act=act); e.forward_debug(ctrl, 0)
    torch.cuda.synchronize()
    fa, fb = a.t["tap_actuator_force"].cpu().numpy(), b.t["tap_actuator_force"].cpu().numpy()
    a.set_state(qpos=q, qvel=np.zeros((4, 1)), act=np.zeros((4, 6))); a.forward_debug(ctrl, 0); torch.cuda.synchronize()
    passive = a.t["tap_actuator_force"].cpu().numpy()
    np.testing.assert_allclose(fb - passive, 0.5 * (fa - passive), rtol=1e-12, atol=1e-9)   # gainprm[:,2]*=0.5, biasprm untouched (base_v0.py:62-67)


def test_walk_random_reset_distribution():
    """WalkEnvV0 reset_type="random" (walk_v0.py:321-337): keyframe 2 or 3 with probability 1/2, N(0, 0.02) noise on every qpos
    coordinate except the height (the reference's quaternion "restore" is a no-op on a view, so the quaternion is noisy too)."""
    import torch
    from myosuite_b200 import vec_env
    n = 2048
    env = vec_env.MyoVecEnv("myoLegWalk-v0", n, auto_reset=False, reset_type="random")
    m = env.mj_model
    env.reset(seed=11); torch.cuda.synchronize()
    q = env.t["qpos"].cpu().numpy(); v = env.t["qvel"].cpu().numpy()
    k2, k3 = m.key_qpos[2], m.key_qpos[3]
    is2 = np.abs(q - k2).sum(1) < np.abs(q - k3).sum(1)
    assert 0.42 < is2.mean() < 0.58
    base = np.where(is2[:, None], k2[None], k3[None]); d = q - base
    assert np.all(d[:, 2] == 0)                                                     # height untouched
    rest = np.delete(d, 2, axis=1)
    assert abs(rest.std() - 0.02) < 0.002 and abs(rest.mean()) < 0.002 and np.abs(rest).max() < 0.12
    assert np.std(d[:, 3:7]) > 0.01                                                 # quaternion noise kept, as in the reference
    np.testing.assert_allclose(v, np.where(is2[:, None], m.key_qvel[2][None], m.key_qvel[3][None]))
    a = torch.zeros(n, env.act_dim, device=env.device)
    obs, rew, done, trunc, _ = env.step(a); torch.cuda.synchronize()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()


def test_fatigue_reset_modes():
    """CumulativeFatigue.reset (fatigue.py:82-99) on the device: default, fatigue_reset_vec, fatigue_reset_random -- also for auto-resets."""
    import torch
    from myosuite_b200 import vec_env
    eid, n = "myoFatiElbowPose1D6MRandom-v0", 256
    env = vec_env.MyoVecEnv(eid, n, seed=1); env.reset(seed=1)
    F = env.t["fatigue"].cpu().numpy(); assert np.all(F[:, 0] == 0) and np.all(F[:, 1] == 1) and np.all(F[:, 2] == 0)
    vec = np.linspace(0.1, 0.6, env.mj_model.na)
    env = vec_env.MyoVecEnv(eid, n, seed=1, fatigue_reset_vec=vec); env.reset(seed=1)
    F = env.t["fatigue"].cpu().numpy()
    np.testing.assert_array_equal(F[:, 2], np.tile(vec, (n, 1))); np.testing.assert_array_equal(F[:, 1], np.tile(1 - vec, (n, 1))); assert np.all(F[:, 0] == 0)
    for _ in range(env.max_episode_steps + 2):                     # run through a TimeLimit auto-reset: the vector is applied again
        env.step(torch.rand(n, env.act_dim, device=env.device) * 2 - 1)
    sc = env.t["step_count"].cpu().numpy(); F = env.t["fatigue"].cpu().numpy()
    assert sc.max() <= 2 and np.all(F[:, 2] > 0.05)                # two steps after the reset MF is still close to the vector (recovery is slow)
    env = vec_env.MyoVecEnv(eid, n, seed=2, fatigue_reset_random=True); env.reset(seed=2)
    F = env.t["fatigue"].cpu().numpy()
    np.testing.assert_allclose(F.sum(1), 1.0, atol=1e-12); assert F.min() >= 0 and F.max() <= 1
    assert 0.4 < F[:, 2].mean() < 0.6 and 0.2 < F[:, 0].mean() < 0.3 and F[:, 0].std() > 0.1          # MF = 1 - u1 ; MA = u1 u2
    with pytest.raises(AssertionError):
        vec_env.MyoVecEnv(eid, 4, fatigue_reset_vec=vec, fatigue_reset_random=True)
    with pytest.raises(AssertionError):
        vec_env.MyoVecEnv(eid, 4, fatigue_reset_vec=vec[:3])


def test_custom_obs_keys():
    """obs_keys kwarg (env_base.py:110,190,458; base_v0.py:33-37 appends "act"): the observation is the concatenation of the requested keys."""
    import torch
    import myosuite_b200 as myo
    from myosuite_b200 import vec_env
    eid, n = "myoHandPoseRandom-v0", 32
    ref = vec_env.MyoVecEnv(eid, n, seed=4); ref.reset(seed=4)
    env = vec_env.MyoVecEnv(eid, n, seed=4, obs_keys=["pose_err", "qpos"]); obs0, _ = env.reset(seed=4)
    m = env.mj_model
    assert env.obs_keys == ["pose_err", "qpos", "act"] and env.obs_dim == 2 * m.nq + m.na and tuple(obs0.shape) == (n, env.obs_dim)
    g = torch.Generator(device="cpu").manual_seed(0)
    for _ in range(3):
        a = (torch.rand(n, env.act_dim, generator=g) * 2 - 1).to(env.device)
        o_ref, r_ref, *_ = ref.step(a); o, r, *_ = env.step(a)
        full = o_ref.cpu().numpy(); nq, nv = m.nq, m.nv
        want = np.concatenate([full[:, nq + nv:2 * nq + nv], full[:, :nq], full[:, 2 * nq + nv:]], axis=1)
        assert np.array_equal(o.cpu().numpy(), want) and torch.equal(r, r_ref)
    assert env.task_info()["solved"].shape[0] == n                      # the info path still reads the full layout
    with pytest.raises(KeyError):
        vec_env.MyoVecEnv(eid, 4, obs_keys=["qpos", "nope"])
    e1 = myo.make("myoElbowPose1D6MRandom-v0", seed=1, obs_keys=["qvel", "pose_err"]).unwrapped
    obs, _ = e1.reset()
    assert e1.obs_keys == ["qvel", "pose_err", "act"] and e1.observation_space.shape == obs.shape == (1 + 1 + 6,)
    obs, *_ = e1.step(np.zeros(6, dtype=np.float32))
    np.testing.assert_array_equal(obs, np.concatenate([np.ravel(e1.obs_dict[k]) for k in e1.obs_keys]))
