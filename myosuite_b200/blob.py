"""Flat "model blob" shared by the C-ABI library and the CPU oracle.

A compiled :class:`~myosuite_b200.mjcf.Model` is packed into one int32 array ``I`` and one float64
array ``D``.  ``I`` starts with a header (magic, version, #dims, #sections), the dims, then a
(kind, offset, length) triple per section; section payloads follow (ints in ``I``, doubles in ``D``).
The section / dim ids are emitted to ``include/myo_blob_layout.h`` by :func:`emit_header`
(``python -m myosuite_b200.blob``), and ``tests/test_blob.py`` checks the header is in sync.

This is the device-side data format too: the library copies ``I``/``D`` to HBM once per model and the
kernels stage the hot sections into shared memory with bulk async copies.
"""
import numpy as np

MAGIC = 0x4D594F42  # 'MYOB'
VERSION = 5

DIMS = ["nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nsite", "ntendon", "nwrap", "nM", "npair", "neq", "nkey",
        "iterations", "ls_iterations"]

# (name, kind, source) ; kind 'i' -> int32 section in I, 'd' -> float64 section in D
RAW_SECTIONS = [
    ("opt", "d"),  # timestep, gravity[3], tolerance, ls_tolerance, meaninertia, impratio
    ("body_parentid", "i"), ("body_rootid", "i"), ("body_weldid", "i"), ("body_jntadr", "i"), ("body_jntnum", "i"),
    ("body_dofadr", "i"), ("body_dofnum", "i"),
    ("body_pos", "d"), ("body_quat", "d"), ("body_ipos", "d"), ("body_iquat", "d"), ("body_mass", "d"),
    ("body_inertia", "d"), ("body_invweight0", "d"),
    ("jnt_type", "i"), ("jnt_qposadr", "i"), ("jnt_dofadr", "i"), ("jnt_bodyid", "i"), ("jnt_limited", "i"),
    ("jnt_pos", "d"), ("jnt_axis", "d"), ("jnt_range", "d"), ("jnt_stiffness", "d"), ("jnt_margin", "d"),
    ("jnt_solref", "d"), ("jnt_solimp", "d"),
    ("dof_bodyid", "i"), ("dof_jntid", "i"), ("dof_parentid", "i"), ("dof_Madr", "i"),
    ("dof_armature", "d"), ("dof_damping", "d"), ("dof_invweight0", "d"),
    ("qpos0", "d"), ("qpos_spring", "d"),
    ("geom_type", "i"), ("geom_bodyid", "i"), ("geom_pos", "d"), ("geom_quat", "d"), ("geom_size", "d"),
    ("site_bodyid", "i"), ("site_pos", "d"),
    ("tendon_adr", "i"), ("tendon_num", "i"), ("wrap_type", "i"), ("wrap_objid", "i"), ("wrap_sidesite", "i"),
    ("actuator_trntype", "i"), ("actuator_trnid", "i"), ("actuator_dyntype", "i"), ("actuator_gaintype", "i"),
    ("actuator_biastype", "i"), ("actuator_ctrllimited", "i"),
    ("actuator_dynprm", "d"), ("actuator_gainprm", "d"), ("actuator_biasprm", "d"), ("actuator_ctrlrange", "d"),
    ("actuator_gear", "d"), ("actuator_lengthrange", "d"), ("actuator_acc0", "d"),
    ("eq_obj1id", "i"), ("eq_obj2id", "i"), ("eq_active0", "i"), ("eq_data", "d"), ("eq_solref", "d"), ("eq_solimp", "d"),
    ("pair_geom1", "i"), ("pair_geom2", "i"), ("pair_dim", "i"),
    ("pair_friction", "d"), ("pair_solref", "d"), ("pair_solimp", "d"), ("pair_margin", "d"), ("pair_gap", "d"),
    ("key_qpos", "d"), ("key_qvel", "d"),
]

# kernel "program" sections (index lists precomputed by myosuite_b200.program); appended after the raw ones
PROGRAM_SECTIONS = [
    ("P_dims", "i"),
    # dynamic bodies (level order)
    ("PB_level_adr", "i"), ("PB_parent", "i"), ("PB_jadr", "i"), ("PB_jnum", "i"), ("PB_model_id", "i"), ("PB_d", "d"),
    ("PD_body", "i"), ("PD_lin", "i"), ("PDOF_d", "d"),
    ("PCH_adr", "i"), ("PCH", "i"), ("PSUB_adr", "i"), ("PSUB", "i"),
    ("PM_i", "i"), ("PM_j", "i"), ("PROW_adr", "i"), ("PROW_col", "i"), ("PROW_idx", "i"),
    # tendons
    ("PPT_body", "i"), ("PPT_xyz", "d"), ("PSP", "i"), ("PWE", "i"), ("PWE_d", "d"),
    ("PT_const", "d"), ("PT_piece_adr", "i"), ("PT_piece", "i"), ("PT_nz_adr", "i"),
    ("PNZ_dof", "i"), ("PNZ_tendon", "i"), ("PNZ_term_adr", "i"), ("PTERM", "i"), ("PCOL_adr", "i"), ("PCOL", "i"),
    ("PA_tendon", "i"), ("PA_d", "d"), ("PA_cls", "i"), ("PAM_d", "d"),
    # collision / constraints
    ("PG_body", "i"), ("PG_type", "i"), ("PG_d", "d"), ("PPAIR", "i"), ("PPAIR_d", "d"), ("PPAIR_tran", "d"), ("PPATH", "i"),
    ("PLIM", "i"), ("PLIM_d", "d"), ("PEQ", "i"), ("PEQ_d", "d"),
    # level-scheduled tree-sparse L'DL
    ("PLV_adr", "i"), ("PLV", "i"), ("PFE_adr", "i"), ("PFE", "i"), ("PFT_adr", "i"), ("PFT", "i"), ("PDS_adr", "i"), ("PDS", "i"),
    # "hot" copy of everything the kernels read: int lists narrowed to int16 (packed two per int32 word) and the double
    # tables, each contiguous, staged into shared memory with one bulk async copy per CTA; HOT_off[sec] = offset or -1
    ("HOT_I16", "i"), ("HOT_D", "d"), ("HOT_off", "i"),
]

# sections the kernels read (everything else in the blob is for the host / the oracle)
KERNEL_RAW = ["jnt_type", "jnt_qposadr", "jnt_dofadr", "dof_Madr", "jnt_pos", "jnt_axis", "jnt_range", "qpos0"]

SECTIONS = RAW_SECTIONS + PROGRAM_SECTIONS
SEC_ID = {name: i for i, (name, _) in enumerate(SECTIONS)}
HDR = 4


def _raw_arrays(m):
    a = {}
    a["opt"] = np.array([m.opt_timestep, *m.opt_gravity, m.opt_tolerance, m.opt_ls_tolerance, m.stat_meaninertia,
                         m.opt_impratio], dtype=np.float64)
    for name, _ in RAW_SECTIONS:
        if name in a:
            continue
        if name == "wrap_sidesite":
            a[name] = m.wrap_prm.astype(np.int32)
        elif name == "actuator_trnid":
            a[name] = m.actuator_trnid[:, 0].copy()
        elif name == "actuator_dynprm":
            a[name] = m.actuator_dynprm[:, :3].copy()
        elif name in ("actuator_gainprm", "actuator_biasprm"):
            a[name] = getattr(m, name)[:, :9].copy()
        elif name == "actuator_gear":
            a[name] = m.actuator_gear[:, 0].copy()
        else:
            a[name] = getattr(m, name)
    return a


# double tables the kernels read straight from HBM / L2 instead of the shared-memory copy: each is touched once per substep (or once per
# stored contact) by lane-parallel, independent loads, so the latency is paid once per phase -- and the 20 KB they free per CTA is what
# lets 14 env-warps of the hand model share one SM with f64-grade constraint rows.  They sit at the END of HOT_D; HOT_off[len(SECTIONS)]
# is the number of doubles that ARE staged.
COLD_D = ["PWE_d", "PPT_xyz", "PPAIR_tran", "PLIM_d", "PAM_d", "PA_d"]


def _hot_pack(arrays):
    kinds = dict(SECTIONS)
    names = KERNEL_RAW + [n for n, _ in PROGRAM_SECTIONS if not n.startswith("HOT_") and n != "P_dims"]
    names = [n for n in names if n not in COLD_D] + [n for n in names if n in COLD_D]
    i16, dd, off = [], [], np.full(len(SECTIONS) + 1, -1, dtype=np.int64)
    ni = nd = 0
    for n in names:
        a = np.asarray(arrays.get(n, np.zeros(0)))
        if kinds[n] == "i":
            v = a.astype(np.int64).ravel()
            if v.size and (v.max() > 32767 or v.min() < -32768):
                raise ValueError("section %s does not fit int16" % n)
            off[SEC_ID[n]] = ni
            i16.append(v.astype(np.int16)); ni += v.size
        else:
            if n in COLD_D and off[len(SECTIONS)] < 0:
                if nd % 2:
                    dd.append(np.zeros(1)); nd += 1                        # staged part: a 16-byte multiple (bulk-copy granularity)
                off[len(SECTIONS)] = nd
            v = a.astype(np.float64).ravel()
            off[SEC_ID[n]] = nd
            dd.append(v); nd += v.size
    I16 = np.concatenate(i16) if i16 else np.zeros(0, np.int16)
    I16 = np.concatenate([I16, np.zeros((-I16.size) % 8, np.int16)])          # 16-byte multiple
    Dh = np.concatenate(dd) if dd else np.zeros(0)
    Dh = np.concatenate([Dh, np.zeros((-Dh.size) % 2)])
    return {"HOT_I16": I16.view(np.int32), "HOT_D": Dh, "HOT_off": off}


def pack(m, program=None):
    """Model (+ optional program dict name->array) -> (I int32[], D float64[])."""
    arrays = _raw_arrays(m)
    if program:
        arrays.update(program)
        arrays.update(_hot_pack(arrays))
    dims = [int({"iterations": m.opt_iterations, "ls_iterations": m.opt_ls_iterations}.get(d, getattr(m, d, 0)))
            for d in DIMS]
    head = HDR + len(DIMS) + 3 * len(SECTIONS)
    ints, dbls, table = [], [], []
    ioff, doff = head, 0
    for name, kind in SECTIONS:
        arr = arrays.get(name)
        if arr is None:
            arr = np.zeros(0)
        if kind == "i":
            v = np.ascontiguousarray(arr, dtype=np.int64).ravel()
            if v.size and (v.max() > 2**31 - 1 or v.min() < -2**31):
                raise ValueError(name)
            v = v.astype(np.int32)
            table += [0, ioff, v.size]
            ints.append(v)
            ioff += v.size
        else:
            v = np.ascontiguousarray(arr, dtype=np.float64).ravel()
            # keep every double section 16-byte aligned (bulk async copies need 16 B granularity)
            pad = (-v.size) % 2
            table += [1, doff, v.size]
            dbls.append(v)
            if pad:
                dbls.append(np.zeros(pad))
            doff += v.size + pad
    I = np.concatenate([np.array([MAGIC, VERSION, len(DIMS), len(SECTIONS)] + dims + table, dtype=np.int32)] + ints)
    pad = (-I.size) % 4
    if pad:
        I = np.concatenate([I, np.zeros(pad, np.int32)])
    D = np.concatenate(dbls) if dbls else np.zeros(0)
    return np.ascontiguousarray(I), np.ascontiguousarray(D)


def section(I, D, name):
    sid = SEC_ID[name]
    base = HDR + int(I[2]) + 3 * sid
    kind, off, n = int(I[base]), int(I[base + 1]), int(I[base + 2])
    return (D if kind else I)[off:off + n]


def emit_header():
    lines = ["/* GENERATED by `python -m myosuite_b200.blob` -- do not edit. Layout of the packed model blob. */",
             "#ifndef MYO_BLOB_LAYOUT_H", "#define MYO_BLOB_LAYOUT_H", "",
             "#define MYO_BLOB_MAGIC 0x%X" % MAGIC, "#define MYO_BLOB_VERSION %d" % VERSION,
             "#define MYO_BLOB_HDR %d" % HDR, "#define MYO_NDIM %d" % len(DIMS), "#define MYO_NSEC %d" % len(SECTIONS), "",
             "enum myo_dim_id {"]
    lines += ["  MYO_DIM_%s = %d," % (d, i) for i, d in enumerate(DIMS)]
    lines += ["};", "", "enum myo_sec_id {"]
    lines += ["  MYO_SEC_%s = %d,%s" % (n, i, "  /* float64 */" if k == "d" else "") for i, (n, k) in enumerate(SECTIONS)]
    lines += ["};", "",
              "#define MYO_DIM(I, id) ((I)[MYO_BLOB_HDR + (id)])",
              "#define MYO_SEC_OFF(I, id) ((I)[MYO_BLOB_HDR + MYO_NDIM + 3 * (id) + 1])",
              "#define MYO_SEC_LEN(I, id) ((I)[MYO_BLOB_HDR + MYO_NDIM + 3 * (id) + 2])",
              "#define MYO_ISEC(I, id) ((I) + MYO_SEC_OFF(I, id))",
              "#define MYO_DSEC(I, D, id) ((D) + MYO_SEC_OFF(I, id))", "", "#endif", ""]
    return "\n".join(lines)


if __name__ == "__main__":
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "myo_blob_layout.h")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write(emit_header())
    print("wrote", out)
