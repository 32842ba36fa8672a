"""bench.py -- env-steps/sec of the env.step hot path (random actions), contract per the round driver.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--env ID] [--envs-per-gpu E] [--impl reference]

Own arm: one "step" = ONE control step (10 physics substeps + obs/reward/done/auto-reset) of every env on
every rank.  `value` = env-steps/s with actions already resident in HBM; `e2e` = the same through
MyoVecEnv.step_host (pinned host action -> H2D, step, D2H of obs/reward/done inside the timed region).
Reference arm (--impl reference): the reference's CPU implementation of the path on the host cores.  MuJoCo
is not installable here (no network, not in /opt/wheelhouse), so this is the CPU oracle port
(oracle/libmyo_oracle.so) -- labelled kind="port".
Each env differs (per-env random state and actions); the working set lives in shared memory and the HBM traffic per
step is the compulsory state/action/obs I/O (8 MB for 4096 hand envs, smaller than L2), so L2 is evicted between
timed steps by a 192 MiB write inside the timed region (config.l2; --no-l2-flush turns it off; it costs < 1 %).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_ENV = "myoHandPoseRandom-v0"       # BASELINE.json target config (configs[2]; 23 dof / 39 muscles), 4096 envs/GPU
# algorithmic bytes per env-step (A_io): action f32 + state f64 in/out + obs f32 + reward/done
A_IO = {"myoElbowPose1D6MRandom-v0": 24 + 2 * (8 * 8) + 36 + 5, "myoHandPoseRandom-v0": 156 + 2 * (85 * 8) + 432 + 5,
        "myoHandObjHoldRandom-v0": 156 + 2 * (98 * 8) + 364 + 5 + 36, "myoFatiLegWalk-v0": 320 + 2 * (149 * 8) + 2 * (240 * 8) + 1612 + 5,
        "myoLegWalk-v0": 320 + 2 * (149 * 8) + 1612 + 5,
        "myoHandReachRandom-v0": 156 + 2 * (85 * 8) + 460 + 5 + 120}        # (+ 15 target coordinates f64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_worker(args):
    """One host core: a single env stepped in a loop (reference protocol:
    /root/reference/benchmarks/mjx_benchmark_baseline.py:8-25 -- gym.make, reset, timeit(env.step(random action)))."""
    env_id, budget_s, seed = args          # budget_s: wall-clock seconds of stepping (a bounded sample; the oracle's rate varies 10x with the contact load)
    import numpy as np
    from myosuite_b200 import assets, blob, vec_env
    from oracle import env_oracle
    from oracle.oracle_py import Oracle
    steps, kw, _ = vec_env.env_spec(env_id)
    m = assets.load(vec_env._MODEL_OF_XML[kw["model_path"]])
    o = Oracle(*blob.pack(m))
    rng = np.random.default_rng(seed)
    q = m.qpos0.copy()
    t0 = time.perf_counter()
    pose = "Pose" in env_id
    s = 0
    while time.perf_counter() - t0 < budget_s:
        if s % steps == 0:
            o.reset()
            if pose:
                for j in range(m.njnt):
                    if m.jnt_type[j] != 0:
                        q[m.jnt_qposadr[j]] = rng.uniform(*m.jnt_range[j])
                o.set(qpos=q)
            elif "Walk" in env_id:
                o.set(qpos=m.key_qpos[2], qvel=m.key_qvel[2])
        env_oracle.env_step(o, rng.uniform(-1, 1, m.nu), 10)
        if pose:
            env_oracle.pose_obs(o.f("qpos"), o.f("qvel"), o.f("act"), q, 0.02)
        else:
            o.forward()          # the reference's extra mj_forward for the observed data (robot.py:607)
        s += 1
    return s, time.perf_counter() - t0


def cpu_baseline(env_id, budget_s, threads=1):
    """(env-steps/s, total steps) of `threads` independent single-env loops of the oracle port, one process per host core, each stepping
    for budget_s seconds of wall clock; the rate is total steps / the slowest worker's loop time (process start-up and model load excluded)."""
    if threads == 1:
        n, t = _cpu_worker((env_id, budget_s, 0))
        return n / t, n
    import multiprocessing as mp
    with mp.get_context("fork").Pool(threads) as pool:
        res = pool.map(_cpu_worker, [(env_id, budget_s, i) for i in range(threads)])
    return sum(r[0] for r in res) / max(r[1] for r in res), sum(r[0] for r in res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--env", default=DEFAULT_ENV)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true", help="do not evict L2 between timed steps (default: a 192 MiB write per step)")
    ap.add_argument("--barrier-mode", type=int, default=0)
    ap.add_argument("--lockstep-groups", type=int, default=0)
    ap.add_argument("--solver-tolerance", type=float, default=0.0, help="Newton stop (scaled gradient); 0 = library default")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    W = max(args.warmup, 3)
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle_py
        oracle_py.build()
        thr = max(1, cores)
        budget = float(os.environ.get("MYO_BENCH_CPU_BUDGET_S", 8.0))      # bounded sample: every host core steps its own env for 8 s
        t0 = time.perf_counter()
        v, total = cpu_baseline(args.env, budget, threads=thr)
        ms = (time.perf_counter() - t0) * 1e3
        per_thread = max(total // thr, 1)
        sample = "%d threads x %.0f s of stepping = %d env-steps of %s (1 env each, random actions, reset every episode)" % (thr, budget, total, args.env)
        print(json.dumps({"impl": "reference", "metric": "env-steps/sec (random actions, 4096 envs/GPU)", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": W, "ms_per_step": ms / max(per_thread, 1), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "%s, frame_skip 10, random actions U[-1,1], reset every episode; CPU arm: one single-env loop of the oracle port per host thread" % args.env},
                          "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": thr, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from myosuite_b200 import build, vec_env
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    build.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.envs_per_gpu
    env = vec_env.MyoVecEnv(args.env, n, device=local_rank, seed=0, env_offset=rank * n, barrier_mode=args.barrier_mode, lockstep_groups=args.lockstep_groups, solver_tolerance=args.solver_tolerance)
    env.reset(seed=0)
    nu = env.act_dim
    gen = torch.Generator(device=env.device).manual_seed(1234 + rank)
    # a ring of pre-generated random action batches (U[-1,1], the action_space.sample() distribution)
    ring = [(torch.rand(n, nu, device=env.device, generator=gen) * 2 - 1) for _ in range(16)]
    ring_host = [r.cpu().pin_memory() for r in ring]

    # L2 (126 MB) is evicted between timed steps by writing a 192 MiB buffer: step k+1 reads the state step k wrote from HBM, not from L2
    flush_buf = None if args.no_l2_flush else torch.empty(192 << 20, dtype=torch.uint8, device=env.device)

    def flush_l2(i):
        if flush_buf is not None:
            flush_buf.fill_(i & 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        env.step(ring[i % 16])
    barrier()
    l0 = env.batch.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        flush_l2(i)
        env.step(ring[i % 16])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.batch.launches - l0
    # end-to-end through the public API with host buffers
    for i in range(3):
        env.step_host(ring_host[i % 16])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        flush_l2(i)
        env.step_host(ring_host[i % 16])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e], device=env.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the one collective of the path: gather finished-episode returns at rollout end
        gathered = torch.empty(world * n, device=env.device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, env.t["last_return"])
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        total = world * n * args.steps
        value, e2e = total / (ms * 1e-3), total / (ms_e2e * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        aio = A_IO.get(args.env, 0)
        per_launch_s = ms * 1e-3 / max(launches, 1)
        achieved = aio * n / per_launch_s / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tf):
            traffic = json.load(open(tf)).get(args.env)
        line = {"metric": "env-steps/sec (random actions, 4096 envs/GPU)", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "b200",
                "config": {"workload": "%s, %d envs/GPU, frame_skip 10, random actions U[-1,1], auto-reset" % (args.env, n),
                           "l2": ("not flushed (--no-l2-flush)" if args.no_l2_flush else "flushed between timed steps: a 192 MiB write per step inside the timed region (inputs, 8 MB of state per step, are smaller than L2)"), "parallelism": "env-sharded x%d, no data-path collective" % world},
                "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": n * nu * 4, "d2h_bytes_per_step": n * (env.obs_dim * 4 + 4 + 1)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "note": "fused step is issue/latency bound, not HBM bound (SURVEY 8d): algorithmic bytes %d B/env-step; peak = %s" % (aio, "measured" if peaks else "fallback")},
                "clocks": clocks}
        if not args.no_cpu_baseline:
            from oracle import oracle_py
            oracle_py.build()
            t0 = time.perf_counter()
            v, nst = cpu_baseline(args.env, float(os.environ.get("MYO_BENCH_CPU_BUDGET_S", 10.0)), threads=1)         # bounded sample: 10 s of single-thread stepping
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                    "sample": "%d env-steps of %s, 1 env, 1 thread (%.1f s)" % (nst, args.env, time.perf_counter() - t0)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
