"""On-device PPO.  CPU: the advantage estimator and the observation normaliser against plain numpy loops.
GPU: a short training run on the elbow pose task must raise the reward (the whole loop -- env kernel, policy, GAE, updates -- on the device)."""
import numpy as np
import pytest


def test_gae_matches_reference_loop():
    import torch
    from myosuite_b200.ppo import gae
    rng = np.random.default_rng(0); T, n, g, l = 12, 5, 0.97, 0.95
    r, v, lv = rng.normal(size=(T, n)), rng.normal(size=(T, n)), rng.normal(size=n)
    d = rng.uniform(size=(T, n)) < 0.2
    adv = np.zeros((T, n))
    for e in range(n):                                        # textbook recursion, one env at a time
        run, nxt = 0.0, lv[e]
        for t in range(T - 1, -1, -1):
            live = 0.0 if d[t, e] else 1.0
            delta = r[t, e] + g * nxt * live - v[t, e]
            run = delta + g * l * live * run
            adv[t, e] = run; nxt = v[t, e]
    a, ret = gae(torch.tensor(r), torch.tensor(v), torch.tensor(d), torch.tensor(lv), g, l)
    np.testing.assert_allclose(a.numpy(), adv, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ret.numpy(), adv + v, rtol=1e-12, atol=1e-12)


def test_running_norm_matches_numpy():
    import torch
    from myosuite_b200.ppo import RunningNorm
    rng = np.random.default_rng(1); chunks = [rng.normal(2.0, 3.0, (int(k), 4)) for k in (7, 50, 3, 200)]
    rn = RunningNorm(4, "cpu")
    rn.mean, rn.var = rn.mean.double(), rn.var.double()
    for c in chunks:
        rn.update(torch.tensor(c))
    allx = np.concatenate(chunks)
    np.testing.assert_allclose(rn.mean.numpy(), allx.mean(0), atol=1e-4)          # (the 1e-4 pseudo-count of the prior)
    np.testing.assert_allclose(rn.var.numpy(), allx.var(0), rtol=1e-3)


@pytest.mark.gpu
def test_ppo_improves_elbow_pose():
    import torch
    from myosuite_b200 import ppo, rollout, vec_env
    env = vec_env.MyoVecEnv("myoElbowPose1D6MRandom-v0", 1024, seed=3)
    agent = ppo.PPO(env, unroll=20, minibatches=8, epochs=4, seed=3)
    hist = agent.train(40 * 20 * 1024)
    first = np.mean([h["reward_per_step"] for h in hist[:3]]); last = np.mean([h["reward_per_step"] for h in hist[-3:]])
    print("PPO elbow: reward/step %.3f -> %.3f over %d env-steps, %.0f env-steps/s incl. updates" % (first, last, hist[-1]["steps"], hist[-1]["steps_per_s"]))
    assert last > first + 0.5 and all(np.isfinite(h["value_loss"]) for h in hist)
    # the trained policy plugs into the batched examine_policy (zero-copy: obs and actions stay on the device)
    ev = vec_env.MyoVecEnv("myoElbowPose1D6MRandom-v0", 64, seed=4, auto_reset=False)
    _, summ = rollout.examine_policy(ev, agent, mode="evaluation", seed=4, keep_obs=False)
    rnd = vec_env.MyoVecEnv("myoElbowPose1D6MRandom-v0", 64, seed=4, auto_reset=False)
    g = torch.Generator(device=rnd.device).manual_seed(0)
    _, base = rollout.examine_policy(rnd, lambda o: torch.rand(o.shape[0], rnd.act_dim, device=o.device, generator=g) * 2 - 1, seed=4, keep_obs=False)
    assert summ["returns"].mean() > base["returns"].mean()
