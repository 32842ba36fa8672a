"""Host-side checks that need no GPU: blob layout, program invariants, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from myosuite_b200 import abi, blob, build, program

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layout_header_in_sync():
    assert open(os.path.join(ROOT, "include", "myo_blob_layout.h")).read() == blob.emit_header()


def test_pack_roundtrip(models):
    for name, m in models.items():
        prog, _ = program.build_program(m)
        I, D = blob.pack(m, prog)
        assert I[0] == blob.MAGIC and I[1] == blob.VERSION
        assert np.array_equal(blob.section(I, D, "body_parentid"), m.body_parentid)
        assert np.allclose(blob.section(I, D, "qpos0"), m.qpos0)
        assert np.array_equal(blob.section(I, D, "PM_i"), prog["PM_i"])
        for sname, kind in blob.SECTIONS:      # 16-byte alignment of every double section (bulk-copy granularity)
            if kind == "d":
                sid = blob.SEC_ID[sname]
                assert int(I[blob.HDR + int(I[2]) + 3 * sid + 1]) % 2 == 0


def test_model_dims_match_survey(models):
    # SURVEY.md Appendix B.1
    exp = {"myoelbow_1dof6muscles": (1, 1, 6, 1), "myohand_pose": (23, 23, 39, 116), "myohand_hold": (30, 29, 39, 137), "myolegs": (35, 34, 80, 351)}
    for name, (nq, nv, nu, nM) in exp.items():
        m = models[name]
        assert (m.nq, m.nv, m.nu, m.nM) == (nq, nv, nu, nM)
        assert m.na == nu
    assert models["myohand_pose"].nwrap == 325 and models["myolegs"].neq == 14 and models["myolegs"].nkey == 4


def test_program_invariants(models):
    for name, m in models.items():
        p, info = program.build_program(m)
        d = p["P_dims"]
        nbd = d[program.PD_NBD]
        assert p["PB_level_adr"][-1] == nbd == len(info["dyn_body_ids"])
        # parents precede children in level order
        for k, par in enumerate(p["PB_parent"]):
            assert par < k
        # every actuated tendon owns exactly one actuator (the kernel writes tfrc without atomics)
        assert len(set(p["PA_tendon"].tolist())) == m.nu
        # each structural non-zero of the moment has at least one term and terms reference valid slots
        adr = p["PNZ_term_adr"]
        assert np.all(np.diff(adr) >= 1)
        t = p["PTERM"].reshape(-1, 3)
        assert np.all(t[:, 0] < d[program.PD_NSP] + 2 * d[program.PD_NWE]) and np.all(np.abs(t[:, 2]) == 1)
        # collision path dofs are those moving exactly one of the two bodies
        assert len(p["PPATH"]) == int(p["PPAIR"][:, 4].sum()) if len(p["PPAIR"]) else True


def _header_functions():
    src = open(os.path.join(ROOT, "include", "myo_b200.h")).read()
    return sorted(set(re.findall(r"\b(myo_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    L = ctypes.CDLL(path)
    names = _header_functions()
    assert set(abi.EXPORTS) == set(names), (set(abi.EXPORTS) ^ set(names))
    for n in names:
        assert hasattr(L, n), n
    assert abi.lib().myo_version() >= 1
    V = abi.lib("f64rows")                      # verification build of the same source (f64 row storage): same ABI
    for n in names:
        assert hasattr(V, n), n
    with pytest.raises(abi.MyoError):
        abi.lib("nope")


def test_no_cpu_fallback_without_gpu(models):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = models["myoelbow_1dof6muscles"]
    prog, _ = program.build_program(m)
    dm = abi.DeviceModel(*blob.pack(m, prog))
    d = dm.dims()
    assert (d.nq, d.nv, d.nu) == (1, 1, 6) and d.smem_bytes_per_env > 0
    cfg = abi.MyoTaskCfg()
    cfg.task, cfg.frame_skip = abi.TASK_POSE, 10
    cfg.reaf_dst = cfg.reaf_src = -1
    with pytest.raises(abi.MyoError, match="no CUDA device"):
        abi.Batch(dm, 0, 4, cfg)
    from myosuite_b200 import vec_env
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vec_env.MyoVecEnv("myoElbowPose1D6MRandom-v0", 4)


def test_blob_rejects_bad_magic(models):
    m = models["myoelbow_1dof6muscles"]
    I, D = blob.pack(m, program.build_program(m)[0])
    I2 = I.copy(); I2[0] = 123
    with pytest.raises(abi.MyoError):
        abi.DeviceModel(I2, D)
    I3, D3 = blob.pack(m)       # no program
    with pytest.raises(abi.MyoError, match="program"):
        abi.DeviceModel(I3, D3)


def test_registry_ids():
    from myosuite_b200 import vec_env
    ids = vec_env.registered_ids()
    for want in ("myoElbowPose1D6MRandom-v0", "myoHandPoseRandom-v0", "myoLegWalk-v0", "myoHandObjHoldRandom-v0",
                 "myoFatiLegWalk-v0", "myoSarcHandPoseRandom-v0", "myoReafHandPoseRandom-v0"):
        assert want in ids
    steps, kw, ep = vec_env.env_spec("myoFatiHandPoseRandom-v0")
    assert steps == 100 and kw["muscle_condition"] == "fatigue" and kw["pose_thd"] == 0.7 and len(kw["target_jnt_range"]) == 23
    steps, kw, ep = vec_env.env_spec("myoElbowPose1D6MRandom-v0")
    assert kw["target_jnt_range"]["r_elbow_flex"] == [0, 2.27] and kw["reset_type"] == "random"


def test_hot_blob_leaves_out_unreachable_lists():
    """program.py: a model whose Newton Hessian is always dense (equalities across the tree) and whose integrator uses the dense
    register Cholesky (8 <= nv <= 36) carries no sparse-LDL schedules; the hand (sparse path reachable) keeps them.  Per-muscle
    parameter classes are 17 doubles wide (only the fields the kernel reads; odd stride)."""
    from myosuite_b200 import assets, program
    legs, _ = program.build_program(assets.load("myolegs"))
    hand, _ = program.build_program(assets.load("myohand_pose"))
    assert len(legs["PFT"]) == 0 and len(legs["PLV"]) == 0 and len(legs["PDS"]) == 0 and len(hand["PFT"]) > 0 and len(hand["PLV"]) > 0
    assert program.PA_STRIDE == 17 and legs["PA_d"].shape[1] == 17 and hand["PA_d"].shape[1] == 17
    m = assets.load("myohand_pose")
    cls = np.asarray(hand["PA_cls"]); row = hand["PA_d"][cls[0]]
    assert np.allclose(row[0:3], m.actuator_dynprm[0, :3]) and row[5] == m.actuator_gainprm[0, 4] and row[8] == m.actuator_gainprm[0, 8]
    assert row[11] == m.actuator_biasprm[0, 5] and row[12] == m.actuator_biasprm[0, 7] and hand["PAM_d"][0, 0] == m.actuator_gainprm[0, 2]


def test_program_invariants_the_kernel_relies_on():
    """Structural promises of program.py that device code depends on without checking:
    * every collision pair lists its dofs in strictly ascending order (the J'WJ assembly maps entry (ei >= ej) to H[di >= dj]);
    * analytic pairs come first, iterative (ellipsoid) pairs after, and P_dims records the split;
    * the number of limit rows that can be active at once bounds the row arrays (one per limited joint unless range < 2 margin)."""
    from myosuite_b200 import assets, program
    for name in ("myohand_pose", "myohand_hold", "myolegs"):
        p, _ = program.build_program(assets.load(name))
        pairs = np.asarray(p["PPAIR"]).reshape(-1, program.PPAIR_ISTRIDE); path = np.asarray(p["PPATH"])
        for q in pairs:
            d = path[q[3]:q[3] + q[4]] >> 1
            assert np.all(np.diff(d) > 0), (name, q)
        ct = pairs[:, 5]; n_an = int(p["P_dims"][program.PD_NPAIR_ANALYTIC])
        assert np.all(ct[:n_an] < program.CT_CAP_ELL) and np.all(ct[n_an:] >= program.CT_CAP_ELL)
        lim = np.asarray(p["PLIM_d"]).reshape(-1, program.PLIM_STRIDE)
        expect = sum(2 if (r[1] - r[0]) < 2 * r[2] else 1 for r in lim)
        assert int(p["P_dims"][program.PD_NLIMROW]) == expect and expect >= len(lim)


def test_dims_contact_capacity(models):
    """myo_model_dims reports the contact / row capacities the kernel will run with (a zero here once silently disabled every contact)."""
    from myosuite_b200 import abi, blob, program
    m = models["myohand_pose"]; prog, _ = program.build_program(m)
    dm = abi.DeviceModel(*blob.pack(m, prog))
    cfg = abi.MyoTaskCfg(); cfg.task = abi.TASK_POSE
    d = dm.dims(cfg)
    assert d.maxcon == 32 and d.maxefc == 23 + 4 * 32 and d.smem_bytes_per_env * 14 + d.reserved[1] + 64 <= 232448      # 14 env-warps per SM: 4096 envs in two rounds
    cfg.maxcon = 48
    assert dm.dims(cfg).maxcon == 48
    cfg.maxcon = 1000
    assert dm.dims(cfg).maxcon == 64
    # the legs model must keep 7 env-warps per SM (2048 envs in two rounds) and the verification build the same ABI numbers
    ml = models["myolegs"]; pl, _ = program.build_program(ml); dl = abi.DeviceModel(*blob.pack(ml, pl)).dims()
    assert dl.smem_bytes_per_env * 7 + dl.reserved[1] + 64 <= 232448
    dv = abi.DeviceModel(*blob.pack(m, prog), variant="f64rows").dims(cfg)
    assert dv.maxcon == 64 and dv.smem_bytes_per_env > dm.dims(cfg).smem_bytes_per_env          # plain f64 rows take more room


def test_tendon_grouping_keeps_lane_rounds_full_and_class_pure(models):
    """program.build_program picks the tendon groups as a subset: minimal number of 32-lane rounds over segments and wrap elements first,
    then minimal number of (round, wrap class) combinations, then balanced scratch."""
    import collections
    for name, want_runs in (("myohand_pose", 5), ("myolegs", 3), ("myotorso", 2), ("myoelbow_1dof6muscles", 2)):
        prog, _ = program.build_program(models[name]); P = [int(x) for x in prog["P_dims"]]
        nsp, nwe = P[3], P[4]; sp_split, we_split = P[26], P[27]
        PWE = np.asarray(prog["PWE"]).reshape(-1, 6)
        rounds = sum((x + 31) // 32 for x in (sp_split, nsp - sp_split, we_split, nwe - we_split))
        assert rounds == (nsp + 31) // 32 + (nwe + 31) // 32, (name, rounds)                      # splitting never costs a lane round
        runs = 0
        for a, b in ((0, we_split), (we_split, nwe)):
            cls = [(int(r[3]), int(r[5] != 0)) for r in PWE[a:b]]
            runs += sum(len(set(cls[i:i + 32])) for i in range(0, len(cls), 32))
        assert runs == want_runs, (name, runs)
        # inside-wrap warm-start slots are a permutation of 0 .. n_inside-1
        slots = sorted(int(r[5]) - 1 for r in PWE if r[5] != 0)
        assert slots == list(range(len(slots)))
