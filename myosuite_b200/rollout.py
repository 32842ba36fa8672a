"""Zero-copy policy rollouts and Trace-compatible logging.

* `MLPPolicy`: the mjrl Gaussian-MLP policy the reference's agents use (agents/baslines_NPG/*/best_policy.pickle; forward pass restated
  from mjrl.utils.fc_network / mjrl.policies.gaussian_mlp), evaluated in torch ON THE ENV'S DEVICE: observations and actions never leave
  HBM between `policy` and `env.step` -- the path the reference reaches with SB3 `make_vec_env` on the CPU (agents/sb3_job_script.py:49)
  and that its MJX benchmark measures (benchmarks/mjx_benchmark_PPO.py).  `get_action(obs)` keeps mjrl's single-env numpy signature.
* `Trace`: the group / dataset container of the reference's logger (logger/grouped_datasets.py:45-135,293-366): create_group,
  create_dataset, append_datum(s), stack, get, items, save / load (pickle; h5py is not a dependency here).
* `examine_policy`: the batched counterpart of MujocoEnv.examine_policy_new (envs/env_base.py:853-969): one "Trial<k>" group per env with the
  datasets time / observations / actions / rewards / done / env_infos, recorded at t = 0 .. T with a NaN action row at the end; all envs
  advance together, the whole history stays on the device until the end of the rollout (one transfer), each trial is cut at its own `done`.
"""
import pickle

import numpy as np


class MLPPolicy:
    def __init__(self, params, device=None):
        import torch
        self.torch, self.device = torch, device
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64, device=device)
        self.W, self.b = [f(w) for w in params["W"]], [f(b) for b in params["b"]]
        self.in_shift, self.in_scale, self.out_shift, self.out_scale = (f(params[k]) for k in ("in_shift", "in_scale", "out_shift", "out_scale"))
        self.log_std = f(params["log_std"])
        self.n, self.m = self.W[0].shape[1], self.W[-1].shape[0]

    def mean(self, obs):
        """obs [..., n] torch tensor on the policy's device -> mean action [..., m] (float64)."""
        x = (obs.double() - self.in_shift) / (self.in_scale + 1e-8)
        for i, (W, b) in enumerate(zip(self.W, self.b)):
            x = x @ W.T + b
            if i < len(self.W) - 1:
                x = self.torch.tanh(x)
        return x * self.out_scale + self.out_shift

    def act(self, obs, mode="exploration", generator=None, clip=True):
        """Batched action on the device.  exploration: mean + exp(log_std) N(0,1); evaluation: the mean.  Clipped to the action space like
        mjrl's GymEnv.step does before env.step."""
        a = self.mean(obs)
        if mode == "exploration":
            a = a + self.torch.exp(self.log_std) * self.torch.randn(a.shape, dtype=a.dtype, device=a.device, generator=generator)
        if clip:
            a = a.clamp(-1.0, 1.0)
        return a.float()

    def get_action(self, observation):
        """mjrl signature: numpy obs [n] -> [action, {"mean", "log_std", "evaluation"}]."""
        torch = self.torch
        o = torch.as_tensor(np.asarray(observation, dtype=np.float64), device=self.device)
        mean = self.mean(o).cpu().numpy()
        noise = np.exp(self.log_std.cpu().numpy()) * np.random.randn(self.m)
        return [mean + noise, {"mean": mean, "log_std": self.log_std.cpu().numpy(), "evaluation": mean}]


class Trace:
    """Minimal logger container with the reference's Trace layout: {name: {group: {dataset: list | stacked array | nested dict}}}."""

    def __init__(self, name):
        self.name, self.root = name, {name: {}}
        self.trace, self.index = self.root[name], 0

    def create_group(self, name):
        self.trace[name] = {}

    def create_dataset(self, group_key, dataset_key, dataset_val):
        self.trace.setdefault(group_key, {})
        if dataset_key in self.trace[group_key]:
            raise KeyError("dataset %s already exists in group %s" % (dataset_key, group_key))
        self.trace[group_key][dataset_key] = [dataset_val]

    def append_datum(self, group_key, dataset_key, dataset_val):
        if dataset_key not in self.trace[group_key]:
            self.create_dataset(group_key, dataset_key, dataset_val)
        else:
            self.trace[group_key][dataset_key].append(dataset_val)

    def append_datums(self, group_key, dataset_key_val):
        for k, v in dataset_key_val.items():
            self.append_datum(group_key, k, v)

    def get(self, group_key, dataset_key=None, dataset_ind=None):
        if dataset_key is None:
            return self.trace[group_key]
        d = self.trace[group_key][dataset_key]
        return d if dataset_ind is None else d[dataset_ind]

    def items(self):
        return self.trace.items()

    def __getitem__(self, k):
        return self.trace[k]

    def __len__(self):
        return len(self.trace)

    @staticmethod
    def _stack(v):
        if isinstance(v, list):
            if v and isinstance(v[0], dict):
                return {k: Trace._stack([x[k] for x in v]) for k in v[0]}
            return np.stack([np.asarray(x) for x in v]) if v else np.zeros(0)
        return v

    def stack(self):
        for g in self.trace.values():
            for k in list(g):
                g[k] = self._stack(g[k])

    def save(self, trace_name, **kwargs):
        with open(trace_name, "wb") as f:
            pickle.dump(self.root, f)

    @staticmethod
    def load(trace_path):
        root = pickle.load(open(trace_path, "rb"))
        name = next(iter(root))
        t = Trace(name); t.root = root; t.trace = root[name]
        return t


def examine_policy(env, policy, horizon=None, mode="exploration", seed=None, generator=None, name=None, keep_obs=True):
    """Roll `policy` through every env of a MyoVecEnv (auto_reset must be off) for one episode each; returns (Trace, summary).
    summary: per-env return, length, solved-step count (success = sum(solved) > 5 as in env_base.evaluate_success)."""
    torch = env.torch
    if env.cfg.auto_reset:
        raise ValueError("examine_policy needs MyoVecEnv(..., auto_reset=False): every env runs exactly one episode")
    T = int(horizon or env.max_episode_steps)
    n = env.num_envs
    obs, _ = env.reset(seed=seed)
    env.refresh_obs()                                    # forward() at t = 0 (env_base.py:899-901)
    H = dict(time=torch.zeros(T + 1, n, dtype=torch.float64, device=env.device), rewards=torch.zeros(T + 1, n, device=env.device),
             done=torch.zeros(T + 1, n, dtype=torch.bool, device=env.device), actions=torch.full((T + 1, n, env.act_dim), float("nan"), device=env.device),
             solved=torch.zeros(T + 1, n, dtype=torch.bool, device=env.device), rwd_sparse=torch.zeros(T + 1, n, dtype=torch.float64, device=env.device))
    if keep_obs:
        H["observations"] = torch.zeros(T + 1, n, env.obs_dim, device=env.device)
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    length = torch.zeros(n, dtype=torch.int64, device=env.device)
    ret = torch.zeros(n, dtype=torch.float64, device=env.device)

    def record(t):
        info = env.task_info()
        H["time"][t], H["rewards"][t], H["done"][t] = env.t["time"], env.t["reward"], env.t["done"].bool()
        if keep_obs:
            H["observations"][t] = env.t["obs"]
        if info is not None:
            H["solved"][t], H["rwd_sparse"][t] = info["solved"], info["rwd_sparse"]
    record(0)
    for t in range(T):
        a = policy.act(env.t["obs"], mode=mode, generator=generator) if hasattr(policy, "act") else policy(env.t["obs"])
        H["actions"][t] = a
        env.step(a)
        record(t + 1)
        ret += torch.where(alive, env.t["reward"].double(), torch.zeros_like(ret))
        length += alive.long()
        alive &= ~(env.t["done"].bool() | env.t["truncated"].bool())
    torch.cuda.current_stream(env.device).synchronize()
    Hc = {k: v.cpu().numpy() for k, v in H.items()}
    L = length.cpu().numpy()
    trace = Trace(name or "%s_rollouts" % env.env_id)
    for e in range(n):
        g = "Trial%d" % e
        trace.create_group(g)
        k = int(L[e]) + 1
        grp = trace.trace[g]
        grp["time"], grp["rewards"], grp["done"], grp["actions"] = Hc["time"][:k, e], Hc["rewards"][:k, e], Hc["done"][:k, e], Hc["actions"][:k, e].copy()
        grp["actions"][k - 1] = np.nan
        if keep_obs:
            grp["observations"] = Hc["observations"][:k, e]
        grp["env_infos"] = {"time": Hc["time"][:k, e], "rwd_dense": Hc["rewards"][:k, e], "rwd_sparse": Hc["rwd_sparse"][:k, e], "solved": Hc["solved"][:k, e], "done": Hc["done"][:k, e]}
    solved_steps = np.array([Hc["solved"][1:int(L[e]) + 1, e].sum() for e in range(n)])
    summary = dict(returns=ret.cpu().numpy(), lengths=L, solved_steps=solved_steps, success_pct=float(100.0 * np.mean(solved_steps > 5)))
    return trace, summary
