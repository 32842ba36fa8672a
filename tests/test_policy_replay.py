"""Trained-policy replay: the strongest anchor to real MuJoCo output that the reference itself holds.

The reference commits NPG policies trained for 1000 iterations on MuJoCo + its own env code, with their logged returns
(agents/baslines_NPG/<env>/*/*/{iterations/best_policy.pickle, logs/log.csv}); tests/golden/npg_policies.py extracts the weights and the
logged statistics into tests/golden/npg_policies.npz.  A trained policy is tuned to the dynamics it saw, so recovering its logged return on
this repo's physics pins the physics where the policy operates.  Protocol: mjrl's sampler (horizon = max_episode_steps, Gaussian noise
exp(log_std), actions clipped to the action space, success = more than 5 solved steps).

What matches and what does not (Feb 2022 logs vs the present reference tree):
  * elbow (model unchanged since): mean / std / max / min of the return agree with the log to < 1 %  -> asserted tightly;
  * hand pose / reach: the hand XML has changed since the logs were written (FDS/FDP wraps at the MCP joints commented out, the whole muscle
    block replaced: simhive/myo_sim/hand/assets/myohand_assets.xml:129-240,540-582), so only the bulk of the trained policies' gain over an
    untrained policy is recovered (pose-random: 90 %, reach-fixed: success 94 % vs 100 %) -> asserted as lower bounds;
  * myoHandObjHoldRandom-v0 (hand + free object: the contact-rich model; all three logged seeds): return -199.5 / -198.5 / -196.4 (std 28.7 /
    26.1 / 27.0, success 1.0 / 0.5 / 0.5 %) against the logged -192.7 / -193.3 / -188.8 (std 26.3 / 24.3 / 23.7, success 1.0 / 0.0 / 1.0 %);
    an untrained policy scores -245 -> 88 % of the trained gain on a drifted hand model, same spread, same success rate -> asserted on the device;
    the Fixed variant's policies (trained to one object pose) keep 60-75 % (two seeds) or nothing (one seed) of their return: documented only;
  * myoHandReachRandom-v0 does not transfer (return 9 +- 150 vs 583 logged): with random targets the episodes end at the far_th test a few
    steps in unless the policy closes the distance at once, which the 2022 muscles did and today's do not.  Documented, not asserted;
  * myoHandPoseFixed-v0 (registration marked "revisit" in the reference) does not transfer at all: the policy drives mcp5 into its limit where
    the 2022 model flexed pm5 -- the little-finger flexors are exactly the tendons whose MCP wraps were removed.  Documented, not asserted.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "devtools")); sys.path.insert(0, os.path.join(HERE, "golden"))

UNTRAINED = {"myoHandPoseRandom-v0": -336.1, "myoHandReachFixed-v0": -20.6, "myoHandObjHoldRandom-v0": -244.7}      # iteration 0 of the same logs (stoc_pol_mean)


@pytest.mark.parametrize("env_id", ["myoElbowPose1D6MFixed-v0", "myoElbowPose1D6MRandom-v0"])
def test_elbow_trained_policy_return_matches_reference_log(env_id):
    import policy_replay
    r = policy_replay.run(env_id, 96, "stochastic", 0, True, procs=4)
    lg = r["logged"]
    lo, hi = min(lg["stoc_pol_mean"], lg["best_stoc_pol_mean"]), max(lg["stoc_pol_mean"], lg["best_stoc_pol_mean"])      # best_policy.pickle is the best iteration's policy
    se = r["std"] / np.sqrt(r["n"])
    assert lo - 4 * se - 2 <= r["mean"] <= hi + 4 * se + 2, (r["mean"], lo, hi)
    assert abs(r["std"] - lg["stoc_pol_std"]) < 0.25 * lg["stoc_pol_std"]
    assert abs(r["max"] - lg["stoc_pol_max"]) < 5 and abs(r["min"] - lg["stoc_pol_min"]) < (20 if "Fixed" in env_id else 45)      # (the minimum over 96 random targets is an extreme-value statistic)
    assert r["success_pct"] == 100.0 == lg["success_percentage"]


def test_hand_trained_policies_keep_most_of_their_gain():
    import policy_replay
    r = policy_replay.run("myoHandPoseRandom-v0", 48, "stochastic", 0, True, procs=8)
    gain_logged = r["logged"]["stoc_pol_mean"] - UNTRAINED["myoHandPoseRandom-v0"]
    assert (r["mean"] - UNTRAINED["myoHandPoseRandom-v0"]) > 0.8 * gain_logged, r
    r = policy_replay.run("myoHandReachFixed-v0", 32, "stochastic", 1, True, procs=8)
    assert r["success_pct"] >= 80.0 and (r["mean"] - UNTRAINED["myoHandReachFixed-v0"]) > 0.6 * (r["logged"]["stoc_pol_mean"] - UNTRAINED["myoHandReachFixed-v0"]), r


@pytest.mark.gpu
def test_trained_policy_replay_on_device_and_trace():
    """The same replay through MyoVecEnv.examine_policy: policy on the device, observations / actions never leave HBM, Trace-compatible log."""
    import torch
    import npg_policies
    from myosuite_b200 import rollout, vec_env
    pols = npg_policies.load_npz()
    for env_id, n, k in (("myoElbowPose1D6MFixed-v0", 1024, 0), ("myoElbowPose1D6MRandom-v0", 1024, 0), ("myoHandPoseRandom-v0", 1024, 0), ("myoHandReachFixed-v0", 512, 1),
                          ("myoHandObjHoldRandom-v0", 1024, 0), ("myoHandObjHoldRandom-v0", 1024, 1), ("myoHandObjHoldRandom-v0", 1024, 2)):
        env = vec_env.MyoVecEnv(env_id, n, auto_reset=False, seed=11)
        pol = rollout.MLPPolicy(pols[env_id][k], device=env.device)
        g = torch.Generator(device=env.device).manual_seed(3)
        trace, s = env.examine_policy(pol, mode="exploration", seed=11, generator=g)
        lg = pols[env_id][k]["logged"]; mean = float(s["returns"].mean())
        print("%s: return %.1f +- %.1f (logged last %.1f / best %.1f), success %.1f %% (logged %.1f)" % (env_id, mean, s["returns"].std() / np.sqrt(n), lg["stoc_pol_mean"], lg["best_stoc_pol_mean"], s["success_pct"], lg["success_percentage"]))
        if "Elbow" in env_id:
            assert min(lg["stoc_pol_mean"], lg["best_stoc_pol_mean"]) - 4 <= mean <= max(lg["stoc_pol_mean"], lg["best_stoc_pol_mean"]) + 4 and s["success_pct"] > 99.0
        elif "PoseRandom" in env_id:
            assert mean - UNTRAINED[env_id] > 0.8 * (lg["stoc_pol_mean"] - UNTRAINED[env_id])
        elif "ObjHold" in env_id:      # contact-rich anchor: most of the trained gain, the logged spread and the logged (near-zero) success rate
            assert mean - UNTRAINED[env_id] > 0.8 * (lg["stoc_pol_mean"] - UNTRAINED[env_id]) and mean < lg["stoc_pol_mean"] + 5
            assert abs(float(s["returns"].std()) - lg["stoc_pol_std"]) < 0.25 * lg["stoc_pol_std"] and s["success_pct"] <= 3.0
        else:
            assert s["success_pct"] >= 80.0
        # Trace layout of the reference's logger: one Trial group per env, T+1 rows, NaN action in the last row
        tr = trace["Trial0"]; L = int(s["lengths"][0])
        assert tr["observations"].shape == (L + 1, env.obs_dim) and tr["actions"].shape == (L + 1, env.act_dim) and np.isnan(tr["actions"][-1]).all()
        assert abs(float(tr["rewards"][1:].sum()) - float(s["returns"][0])) < 1e-3 and set(tr["env_infos"]) >= {"time", "rwd_dense", "rwd_sparse", "solved", "done"}
    # single-env facade with the reference's examine_policy_new signature (mjrl policy.get_action)
    import myosuite_b200 as myo
    e1 = myo.make("myoElbowPose1D6MFixed-v0", seed=1)
    tr = e1.examine_policy_new(rollout.MLPPolicy(pols["myoElbowPose1D6MFixed-v0"][0], device=e1.vec.device), horizon=20, num_episodes=2, mode="evaluation")
    assert len(tr) == 2 and tr["Trial1"]["observations"].shape == (21, 9) and tr["Trial0"]["env_infos"]["rwd_dense"].shape == (21,)
