"""Kernel "program" builder: turns a compiled Model into the flat index lists the CUDA kernels execute.

The kernels run one env per warp; every phase is a loop ``for (i = lane; i < n; i += 32)`` over a
list built here, so that all data-dependent structure (tree levels, which dofs lie between two
bodies, which tendon segments are compile-time constants, which wrap is inside/outside, which
geom pairs survive the static collision filters) is resolved once per model on the host.

Formulation notes (deliberately different from the oracle's MuJoCo-style formulation):
  * only DYNAMIC bodies (those below a joint) are simulated; static bodies are folded into constants;
  * spatial quantities are taken about the world origin, not the subtree COM;
  * a tendon segment whose two end points ride on the same body has constant length and no moment:
    it is summed into ``PT_const`` at build time;
  * the tendon moment is assembled per structural non-zero (tendon, dof) from "terms"
    (segment direction . (axis x (point - anchor))) for dofs lying between the segment's bodies.
"""
import numpy as np

from . import mjcf
from . import mjmath as mm

# P_dims slots
(PD_NBD, PD_NLEVEL, PD_NPT, PD_NSP, PD_NWE, PD_NTA, PD_NNZ, PD_NTERM, PD_NLIM, PD_NEQ, PD_NPAIR, PD_NGC, PD_MAXPATH,
 PD_MAXCHAIN, PD_NSUB, PD_NROW, PD_NCOL, PD_NPIECE, PD_NWE_SPH_OUT, PD_NWE_SPH_IN, PD_NWE_CYL_OUT, PD_NWE_CYL_IN,
 PD_NDEPTH, PD_EQ_TREE, PD_NPAIR_ANALYTIC, PD_NLIMROW, PD_SPLIT_SP, PD_SPLIT_WE, PD_SPLIT_TA, PD_SPLIT_NZ) = range(30)
NPDIM = 30

PB_STRIDE, PWE_STRIDE, PA_STRIDE, PG_STRIDE, PPAIR_STRIDE, PLIM_STRIDE, PEQ_STRIDE = 22, 16, 17, 16, 12, 12, 16
PAM_STRIDE, PPAIR_ISTRIDE = 6, 8

# collision function ids
CT_NONE, CT_CAP_CAP, CT_SPH_SPH, CT_SPH_CAP, CT_PLANE_SPH, CT_PLANE_CAP, CT_PLANE_ELL, CT_CAP_ELL, CT_ELL_ELL = range(9)


def _kbimp(solref, solimp, timestep):
    """Clamp solimp like MuJoCo's getsolparam and precompute K, B (refsafe) -> (K, B, solimp[5])."""
    si = np.array(solimp, dtype=np.float64).copy()
    si[0] = np.clip(si[0], 0.0001, 0.9999); si[1] = np.clip(si[1], 0.0001, 0.9999); si[2] = max(0.0, si[2])
    si[3] = np.clip(si[3], 0.0001, 0.9999); si[4] = max(1.0, si[4])
    if solref[0] <= 0:
        raise mjcf.MJCFError("direct (negative) solref not supported")
    tc = max(solref[0], 2 * timestep)
    K = 1.0 / max(1e-15, si[1] ** 2 * tc ** 2 * solref[1] ** 2)
    B = 2.0 / max(1e-15, si[1] * tc)
    return K, B, si


# Collision pairs whose geom-type combination has no device collider are an ERROR unless listed here by (model name of geom1, geom2).
# myohand_hold: the free object ellipsoid vs the scene's static pedestal cylinder.  The pair can only act after the object has fallen
# 1.4 m; ObjHold terminates ("drop") at 0.3 m from the goal (obj_hold_v0.py:100), so it never produces a contact inside an episode.
# myotorso: the trunk / neck / head collision geoms (all above z = 1.3 m on a pelvis fixed at z ~ 0.9 m; the spine chain is ~0.7 m long) vs the
# scene's floor disk (top at z = 0.015 m): geometrically out of reach, but the bounding-sphere reach test of the compiler cannot prove it
# for a 1.05 m-radius cylinder.
ALLOWED_UNSUPPORTED_PAIRS = {("object", "<static cylinder>"), ("<static cylinder>", "object"),        # unnamed static geoms are labelled <static TYPE>
                             ("<capsule on torso>", "<static cylinder>"), ("hat_cervical_coll", "<static cylinder>"), ("hat_jaw_coll2", "<static cylinder>"),
                             ("hat_skull_coll", "<static cylinder>")}


def build_program(m, allow_unsupported=ALLOWED_UNSUPPORTED_PAIRS):
    nb = m.nbody
    dyn_ids = [b for b in range(1, nb) if m.body_weldid[b] != 0]
    kin0 = mjcf.kinematics(m, m.qpos0)
    # ---- levels (depth among dynamic bodies)
    depth = {}
    for b in dyn_ids:
        p = m.body_parentid[b]
        depth[b] = depth[p] + 1 if p in depth else 0
    order = sorted(dyn_ids, key=lambda b: (depth[b], b))
    idx = {b: k for k, b in enumerate(order)}          # model body id -> dyn index
    nbd = len(order)
    nlevel = (max(depth.values()) + 1) if order else 0
    level_adr = [0] * (nlevel + 1)
    for b in order:
        level_adr[depth[b] + 1] += 1
    level_adr = np.cumsum(level_adr).tolist()

    def bidx(b):
        return idx.get(b, -1)

    PB_parent, PB_jadr, PB_jnum, PB_d = [], [], [], np.zeros((nbd, PB_STRIDE))
    for k, b in enumerate(order):
        p = m.body_parentid[b]
        PB_parent.append(bidx(p))
        PB_jadr.append(int(m.body_jntadr[b])); PB_jnum.append(int(m.body_jntnum[b]))
        if p in idx:
            pos, quat = m.body_pos[b], m.body_quat[b]
        else:   # static parent: fold its world pose in
            pos = kin0["xpos"][p] + kin0["xmat"][p] @ m.body_pos[b]
            quat = mm.quat_normalize(mm.quat_mul(kin0["xquat"][p], m.body_quat[b]))
        R = mm.quat2mat(m.body_iquat[b])
        Il = R @ np.diag(m.body_inertia[b]) @ R.T
        PB_d[k, 0:3], PB_d[k, 3:12], PB_d[k, 12:15], PB_d[k, 15] = pos, mm.quat2mat(quat).ravel(), m.body_ipos[b], m.body_mass[b]
        PB_d[k, 16:22] = [Il[0, 0], Il[1, 1], Il[2, 2], Il[0, 1], Il[0, 2], Il[1, 2]]
        for j in range(m.body_jntadr[b], m.body_jntadr[b] + m.body_jntnum[b]):
            if m.jnt_type[j] == mjcf.JNT_FREE and (p in idx or m.body_jntnum[b] != 1):
                raise mjcf.MJCFError("free joints must be alone on a child of a static body")
    # ---- dofs
    nv = m.nv
    PD_body = [idx[int(b)] for b in m.dof_bodyid]
    PD_lin = []
    for d in range(nv):
        j = m.dof_jntid[d]
        t = m.jnt_type[j]
        PD_lin.append(1 if (t == mjcf.JNT_SLIDE or (t == mjcf.JNT_FREE and d - m.jnt_dofadr[j] < 3)) else 0)
    if np.any(m.jnt_stiffness != 0):
        raise mjcf.MJCFError("joint stiffness not supported (all hot-path models use 0)")
    PDOF_d = np.stack([m.dof_armature, m.dof_damping], axis=1) if nv else np.zeros((0, 2))
    chains = {b: mjcf.dof_chain(m, b) for b in order}
    PCH_adr, PCH = [0], []
    for b in order:
        for d in chains[b]:
            j = m.dof_jntid[d]
            flag = 1 if (m.jnt_type[j] == mjcf.JNT_FREE and d - m.jnt_dofadr[j] >= 4) else 0
            PCH.append(d * 2 + flag)
        PCH_adr.append(len(PCH))
    maxchain = max([len(c) for c in chains.values()] + [0])
    # subtrees among dynamic bodies (self included)
    subtree = {b: [b] for b in order}
    for b in reversed(order):
        p = m.body_parentid[b]
        if p in subtree:
            subtree[p] = subtree[p] + subtree[b]
    PSUB_adr, PSUB = [0], []
    for b in order:
        PSUB += [idx[x] for x in sorted(subtree[b])]
        PSUB_adr.append(len(PSUB))
    # mass-matrix entries (same order as qM: row i -> i, parent(i), ...)
    PM_i, PM_j = [], []
    for i in range(nv):
        j = i
        while j >= 0:
            PM_i.append(i); PM_j.append(j)
            j = m.dof_parentid[j]
    assert len(PM_i) == m.nM
    rows = [[] for _ in range(nv)]
    for e, (i, j) in enumerate(zip(PM_i, PM_j)):
        rows[i].append((j, e))
        if i != j:
            rows[j].append((i, e))
    PROW_adr, PROW_col, PROW_idx = [0], [], []
    for i in range(nv):
        for c, e in sorted(rows[i]):
            PROW_col.append(c); PROW_idx.append(e)
        PROW_adr.append(len(PROW_col))

    # ---- level-scheduled tree-sparse L'DL (left-looking): dofs by depth; per entry the descendant products to gather
    dpar = m.dof_parentid
    ddepth = np.zeros(nv, dtype=np.int64)
    for d in range(nv):
        ddepth[d] = 0 if dpar[d] < 0 else ddepth[dpar[d]] + 1
    ndepth = int(ddepth.max()) + 1 if nv else 0
    anc = {d: [] for d in range(nv)}            # proper ancestors, nearest first
    for d in range(nv):
        j = dpar[d]
        while j >= 0:
            anc[d].append(int(j)); j = dpar[j]
    eidx = {}                                   # (i, j) -> index in the qM layout (j ancestor-or-self of i)
    for i in range(nv):
        eidx[(i, i)] = int(m.dof_Madr[i])
        for c, j in enumerate(anc[i]):
            eidx[(i, j)] = int(m.dof_Madr[i]) + c + 1
    desc = {d: [] for d in range(nv)}
    for i in range(nv):
        for j in anc[i]:
            desc[j].append(i)
    PLV_adr, PLV = [0], []
    for dep in range(ndepth):
        PLV += [d for d in range(nv) if ddepth[d] == dep]
        PLV_adr.append(len(PLV))
    PFE_adr, PFE, PFT_adr, PFT = [0], [], [0], []
    for dep in range(ndepth):                   # stored ascending; the kernel walks levels deepest-first
        for k in PLV[PLV_adr[dep]:PLV_adr[dep + 1]]:
            for j in [k] + anc[k]:
                PFE.append(eidx[(k, j)])
                for d_ in desc[k]:
                    PFT += [eidx[(d_, k)], eidx[(d_, j)], d_]
                PFT_adr.append(len(PFT) // 3)
        PFE_adr.append(len(PFE))
    PDS_adr, PDS = [0], []
    for j in range(nv):
        for i in desc[j]:
            PDS += [i, eidx[(i, j)]]
        PDS_adr.append(len(PDS) // 2)

    def moves(d, b):
        """does dof d move model body b?"""
        return d in chains.get(b, ())

    # ---- tendon program
    PPT_body, PPT_xyz, pt_cache = [], [], {}

    def point_ref(body, local):
        """(body model id, local pos) -> PT index; static bodies are folded to world coordinates."""
        if body in idx:
            key = (idx[body], tuple(np.round(local, 15)))
            xyz = np.asarray(local, dtype=np.float64)
            bi = idx[body]
        else:
            xyz = kin0["xpos"][body] + kin0["xmat"][body] @ np.asarray(local, dtype=np.float64)
            key = (-1, tuple(np.round(xyz, 15)))
            bi = -1
        if key not in pt_cache:
            pt_cache[key] = len(PPT_body)
            PPT_body.append(bi); PPT_xyz.append(xyz)
        return pt_cache[key]

    def site_ref(s):
        return point_ref(int(m.site_bodyid[s]), m.site_pos[s])

    act_tendons = sorted(set(int(t) for t in m.actuator_trnid[:, 0])) if m.nu else []
    if m.nu and not np.all(m.actuator_trntype == mjcf.TRN_TENDON):
        raise mjcf.MJCFError("only tendon transmissions are supported")
    # Two passes over the tendons (group A = the first `split` program tendons, B = the rest): the unit vectors and piece lengths of a pass are
    # dead once its moments and lengths are formed, so the kernel's scratch holds one group at a time.  The program's tendon ORDER is internal
    # (actuators reach their tendon through PA_tendon), so the groups are chosen as a subset: first keep the number of 32-lane rounds over the
    # straight segments and the wrap elements minimal (a group of 33 wrap elements costs a whole extra round of the wrap code for one lane),
    # then balance the two groups' scratch need.
    def _count(t):
        """(straight segments, wrap elements, bit mask of the wrap classes {sphere, cylinder} x {outside, inside}) of tendon t"""
        adr, num = int(m.tendon_adr[t]), int(m.tendon_num[t]); ns = nw = cls = 0; j = 0; ccount = [0, 0, 0, 0]
        while j < num - 1:
            if m.wrap_type[adr + j + 1] == mjcf.WRAP_SITE:
                ns += 0 if _same_rigid(m, int(m.site_bodyid[int(m.wrap_objid[adr + j])]), int(m.site_bodyid[int(m.wrap_objid[adr + j + 1])])) else 1; j += 1
            else:
                g_, ss_ = int(m.wrap_objid[adr + j + 1]), int(m.wrap_prm[adr + j + 1]); inside_ = 0
                if ss_ >= 0:
                    bg_ = int(m.geom_bodyid[g_])
                    inside_ = int(np.linalg.norm(_site_world(m, kin0, ss_) - (kin0["xpos"][bg_] + kin0["xmat"][bg_] @ m.geom_pos[g_])) < m.geom_size[g_][0])
                c_ = 2 * (0 if m.wrap_type[adr + j + 1] == mjcf.WRAP_SPHERE else 1) + inside_
                cls |= 1 << c_; ccount[c_] += 1
                nw += 1; j += 2
        return ns, nw, cls, tuple(ccount)
    _cnt = {t: _count(t) for t in act_tendons}
    _need = lambda ns, nw: 3 * (ns + 2 * nw) + ns + nw
    # reachable (ns_A, nw_A, classes_A, classes_B) -> bit mask of one subset A reaching it.  A lane round of wrap elements executes every
    # class present in it one after the other (the four wrap variants diverge), so after the round count the number of classes per group
    # is minimised: hand model 4 + 4 -> 4 + 1 (one group is all cylinder-outside wraps).
    _reach = {(0, 0, 0, 0): 0}
    for k_, t in enumerate(act_tendons):
        nxt = {}
        for (ns_, nw_, ca, cb), msk in _reach.items():
            nxt.setdefault((ns_ + _cnt[t][0], nw_ + _cnt[t][1], ca | _cnt[t][2], cb), msk | (1 << k_))       # t joins A
            nxt.setdefault((ns_, nw_, ca, cb | _cnt[t][2]), msk)                                             # t stays in B
        _reach = nxt
    _NS, _NW = sum(c[0] for c in _cnt.values()), sum(c[1] for c in _cnt.values())
    _r32 = lambda x: (x + 31) // 32

    def _class_runs(msk):
        """number of (lane round, wrap class) combinations the two groups execute: elements are sorted by (inside first, type) inside a group
        (the order of we_order below) and dealt to rounds of 32 lanes"""
        tot = 0
        for grp_ in (1, 0):
            cc = [0, 0, 0, 0]
            for k_, t in enumerate(act_tendons):
                if ((msk >> k_) & 1) == grp_:
                    for c_ in range(4):
                        cc[c_] += _cnt[t][3][c_]
            seq = [c_ for c_ in (1, 3, 0, 2) for _ in range(cc[c_])]           # inside sphere, inside cylinder, outside sphere, outside cylinder
            tot += sum(len(set(seq[i:i + 32])) for i in range(0, len(seq), 32))
        return tot
    _best = min(_reach, key=lambda q: (_r32(q[0]) + _r32(_NS - q[0]) + _r32(q[1]) + _r32(_NW - q[1]), _class_runs(_reach[q]),
                                       max(_need(q[0], q[1]), _need(_NS - q[0], _NW - q[1])), q))
    _msk = _reach[_best]
    act_tendons = [t for k_, t in enumerate(act_tendons) if (_msk >> k_) & 1] + [t for k_, t in enumerate(act_tendons) if not (_msk >> k_) & 1]
    split = bin(_msk).count("1")
    ta_index = {t: k for k, t in enumerate(act_tendons)}
    sp_list, we_list, sp_ta = [], [], []      # runtime pieces (sp_ta: tendon of each straight segment)
    T_const, T_pieces = [], []     # per active tendon
    # term bookkeeping: terms[(ta, dof)] = list of (ukind, uidx, ptcode, sign)
    terms = {}

    def add_terms(ta, ba, bb, uref, pa_code, pb_code):
        """straight piece from point a (on body ba) to point b (on bb) with unit-vector slot `uref`."""
        ds = set(chains.get(ba, ())) ^ set(chains.get(bb, ()))
        for d in ds:
            if moves(d, bb):
                terms.setdefault((ta, d), []).append((uref, pb_code, +1))
            else:
                terms.setdefault((ta, d), []).append((uref, pa_code, -1))

    for t in act_tendons:
        ta = ta_index[t]
        adr, num = int(m.tendon_adr[t]), int(m.tendon_num[t])
        const, pieces = 0.0, []
        j = 0
        while j < num - 1:
            t1 = m.wrap_type[adr + j + 1]
            s0 = int(m.wrap_objid[adr + j])
            b0 = int(m.site_bodyid[s0])
            if t1 == mjcf.WRAP_SITE:
                s1 = int(m.wrap_objid[adr + j + 1])
                b1 = int(m.site_bodyid[s1])
                if _same_rigid(m, b0, b1):
                    const += float(np.linalg.norm(_site_world(m, kin0, s1) - _site_world(m, kin0, s0)))
                else:
                    pa, pb = site_ref(s0), site_ref(s1)
                    k = len(sp_list)
                    sp_list.append((pa, pb)); sp_ta.append(ta)
                    pieces.append(("S", k))
                    add_terms(ta, b0, b1, ("S", k), pa, pb)
                j += 1
            else:
                g = int(m.wrap_objid[adr + j + 1]); s1 = int(m.wrap_objid[adr + j + 2]); ss = int(m.wrap_prm[adr + j + 1])
                b1, bg = int(m.site_bodyid[s1]), int(m.geom_bodyid[g])
                pa, pb = site_ref(s0), site_ref(s1)
                typ = 0 if m.wrap_type[adr + j + 1] == mjcf.WRAP_SPHERE else 1
                inside = 0
                side = -1
                s_local = np.zeros(3)
                if ss >= 0:
                    if not _same_rigid(m, int(m.site_bodyid[ss]), bg):
                        raise mjcf.MJCFError("side site must ride on the wrap geom's body")
                    sw, gw = _site_world(m, kin0, ss), kin0["xpos"][bg] + kin0["xmat"][bg] @ m.geom_pos[g]
                    inside = int(np.linalg.norm(sw - gw) < m.geom_size[g][0])
                    side = site_ref(ss)
                    s_local = (kin0["xmat"][bg] @ mm.quat2mat(m.geom_quat[g])).T @ (sw - gw)
                if bg in idx:
                    gpos, gmat, gb = m.geom_pos[g], mm.quat2mat(m.geom_quat[g]), idx[bg]
                else:
                    gpos = kin0["xpos"][bg] + kin0["xmat"][bg] @ m.geom_pos[g]
                    gmat = kin0["xmat"][bg] @ mm.quat2mat(m.geom_quat[g]); gb = -1
                k = len(we_list)
                we_list.append(dict(pa=pa, pb=pb, gb=gb, typ=typ, side=side, inside=inside, gpos=gpos, gmat=gmat,
                                    r=float(m.geom_size[g][0]), ta=ta, b0=b0, b1=b1, bg=bg, s_local=s_local))
                pieces.append(("W", k))
                j += 2
        T_const.append(const); T_pieces.append(pieces)
    nta_all = len(act_tendons)
    grp = lambda ta_: 0 if ta_ < split else 1
    # within a group: wrap elements by (inside, type) so that a warp round is branch-uniform
    we_order = sorted(range(len(we_list)), key=lambda k: (grp(we_list[k]["ta"]), -we_list[k]["inside"], we_list[k]["typ"], k))
    we_new = {old: new for new, old in enumerate(we_order)}
    we_sorted = [we_list[k] for k in we_order]
    nsp, nwe = len(sp_list), len(we_sorted)
    sp0 = [0, sum(1 for t_ in sp_ta if t_ < split), nsp]                      # straight segments are generated in tendon order: already grouped
    we0 = [0, sum(1 for w_ in we_sorted if w_["ta"] < split), nwe]
    nspg = [sp0[1] - sp0[0], sp0[2] - sp0[1]]

    def u_local(ta_, kind, idx):
        """index of a unit-vector slot / piece inside its group's scratch: segments first, then two slots (one length) per wrap element"""
        g_ = grp(ta_)
        return idx - sp0[g_] if kind == "S" else nspg[g_] + (idx - 2 * we0[g_])

    def pl_local(ta_, kind, idx):
        g_ = grp(ta_)
        return idx - sp0[g_] if kind == "S" else nspg[g_] + (idx - we0[g_])
    wz_slot, nwz_ = {}, 0
    for new_, w_ in enumerate(we_sorted):
        if w_["inside"]:
            wz_slot[new_] = nwz_; nwz_ += 1
    for new, w in enumerate(we_sorted):
        # piece A: pa (b0) -> w0 (bg), unit slot nsp+2*new ; piece B: w1 (bg) -> pb (b1), unit slot nsp+2*new+1
        # The moment of a straight piece is  u . (axis x (p - anchor))  for ANY point p on the piece's line (p may slide along u), so the
        # tangent point riding on the wrap geom's body is replaced by the piece's site: no wrap points are stored at run time.
        add_terms(w["ta"], w["b0"], w["bg"], ("W", 2 * new), w["pa"], w["pa"])
        add_terms(w["ta"], w["bg"], w["b1"], ("W", 2 * new + 1), w["pb"], w["pb"])
    counts = [0, 0, 0, 0]
    for w in we_sorted:
        counts[2 * w["typ"] + w["inside"]] += 1
    PT_piece_adr, PT_piece = [0], []
    for ta_, pieces in enumerate(T_pieces):
        for kind, k in pieces:
            PT_piece.append(pl_local(ta_, kind, k if kind == "S" else we_new[k]))
        PT_piece_adr.append(len(PT_piece))
    # non-zeros sorted by (tendon, dof)
    keys = sorted(terms.keys())
    PNZ_dof, PNZ_tendon, PNZ_term_adr, PTERM = [], [], [0], []
    PT_nz_adr = [0] * (len(act_tendons) + 1)
    for (ta, d) in keys:
        PNZ_dof.append(d); PNZ_tendon.append(ta); PT_nz_adr[ta + 1] += 1
        for (ukind, uk), ptcode, sign in terms[(ta, d)]:
            PTERM += [u_local(ta, ukind, uk), ptcode, sign]
        PNZ_term_adr.append(len(PTERM) // 3)
    PT_nz_adr = np.cumsum(PT_nz_adr).tolist()
    cols = [[] for _ in range(nv)]
    for k, d in enumerate(PNZ_dof):
        cols[d].append(k)
    PCOL_adr, PCOL = [0], []
    for d in range(nv):
        PCOL += cols[d]; PCOL_adr.append(len(PCOL))
    # [5]: 0 = outside wrap, else 1 + slot of this element's warm-started inverse-wrap root
    PWE = np.array([[w["pa"], w["pb"], w["gb"], w["typ"], w["side"], (1 + wz_slot[k_]) if w["inside"] else 0] for k_, w in enumerate(we_sorted)], dtype=np.int32).reshape(-1, 6)
    PWE_d = np.zeros((nwe, PWE_STRIDE))
    for k, w in enumerate(we_sorted):
        PWE_d[k, 0:3], PWE_d[k, 3:12], PWE_d[k, 12] = w["gpos"], np.asarray(w["gmat"]).ravel(), w["r"]
        PWE_d[k, 13:16] = w["s_local"]
    # actuators
    PA_tendon = [ta_index[int(t)] for t in m.actuator_trnid[:, 0]] if m.nu else []
    if len(set(PA_tendon)) != m.nu:
        raise mjcf.MJCFError("each actuated tendon must carry exactly one actuator")
    # per-muscle record (peak forces, lengthrange, gear) + shared parameter classes holding only what the kernel reads:
    # [0:3] dynprm (tau_act, tau_deact, tausmooth) | [3:9] gainprm range0, range1, lmin, lmax, vmax, fvmax | [9:13] biasprm range0, range1, lmax, fpmax
    # | [13:15] ctrlrange | [15] ctrllimited | [16] pad (odd stride: a lane per muscle reads its row without bank conflicts)
    PAM_d, PA_cls, cls_rows, cls_index = np.zeros((m.nu, PAM_STRIDE)), [], [], {}
    for i in range(m.nu):
        if not (m.actuator_dyntype[i] == mjcf.DYN_MUSCLE and m.actuator_gaintype[i] == mjcf.GAIN_MUSCLE
                and m.actuator_biastype[i] == mjcf.BIAS_MUSCLE):
            raise mjcf.MJCFError("only muscle actuators are supported on the device path")
        g, bp = m.actuator_gainprm[i], m.actuator_biasprm[i]
        PAM_d[i] = [g[2], bp[2], m.actuator_lengthrange[i, 0], m.actuator_lengthrange[i, 1], m.actuator_gear[i, 0], 0.0]   # gain force, bias force, lengthrange, gear
        row = np.zeros(PA_STRIDE)
        row[0:3] = m.actuator_dynprm[i, :3]
        row[3:9] = [g[0], g[1], g[4], g[5], g[6], g[8]]
        row[9:13] = [bp[0], bp[1], bp[5], bp[7]]
        row[13:15], row[15] = m.actuator_ctrlrange[i], float(m.actuator_ctrllimited[i])
        key = row.tobytes()
        if key not in cls_index:
            cls_index[key] = len(cls_rows); cls_rows.append(row)
        PA_cls.append(cls_index[key])

    # ---- collision geoms + pairs
    h = m.opt_timestep
    gmap, PG_body, PG_type, PG_d = {}, [], [], []

    def geom_ref(g):
        if g not in gmap:
            b = int(m.geom_bodyid[g])
            if b in idx:
                pos, mat, bi = m.geom_pos[g], mm.quat2mat(m.geom_quat[g]), idx[b]
            else:
                pos = kin0["xpos"][b] + kin0["xmat"][b] @ m.geom_pos[g]
                mat = kin0["xmat"][b] @ mm.quat2mat(m.geom_quat[g]); bi = -1
            gmap[g] = len(PG_body)
            PG_body.append(bi); PG_type.append(int(m.geom_type[g]))
            row = np.zeros(PG_STRIDE); row[0:3], row[3:12], row[12:15] = pos, np.asarray(mat).ravel(), m.geom_size[g]
            PG_d.append(row)
        return gmap[g]

    ctype_of = {(mjcf.GEOM_CAPSULE, mjcf.GEOM_CAPSULE): CT_CAP_CAP, (mjcf.GEOM_SPHERE, mjcf.GEOM_SPHERE): CT_SPH_SPH,
                (mjcf.GEOM_SPHERE, mjcf.GEOM_CAPSULE): CT_SPH_CAP, (mjcf.GEOM_PLANE, mjcf.GEOM_SPHERE): CT_PLANE_SPH,
                (mjcf.GEOM_PLANE, mjcf.GEOM_CAPSULE): CT_PLANE_CAP, (mjcf.GEOM_PLANE, mjcf.GEOM_ELLIPSOID): CT_PLANE_ELL,
                (mjcf.GEOM_CAPSULE, mjcf.GEOM_ELLIPSOID): CT_CAP_ELL, (mjcf.GEOM_ELLIPSOID, mjcf.GEOM_ELLIPSOID): CT_ELL_ELL}
    PPAIR, PPAIR_d, PPATH, PPAIR_tran, pcls_index = [], [], [], [], {}
    maxpath = 0
    pair_model_index = []
    def _ct(p):
        return ctype_of.get((int(m.geom_type[int(m.pair_geom1[p])]), int(m.geom_type[int(m.pair_geom2[p])])), CT_NONE)
    dropped_pairs = []
    for p in range(m.npair):
        if _ct(p) == CT_NONE:
            tname = {mjcf.GEOM_PLANE: "plane", mjcf.GEOM_SPHERE: "sphere", mjcf.GEOM_CAPSULE: "capsule", mjcf.GEOM_ELLIPSOID: "ellipsoid"}

            def _label(g):
                nm = m.id2name("geom", g)
                if nm and not nm.startswith("geom"):
                    return nm
                kind = tname.get(int(m.geom_type[g]), {5: "cylinder", 6: "box", 7: "mesh"}.get(int(m.geom_type[g]), "type%d" % int(m.geom_type[g])))
                return "<static %s>" % kind if m.body_weldid[int(m.geom_bodyid[g])] == 0 else "<%s on %s>" % (kind, m.id2name("body", int(m.geom_bodyid[g])))
            names = tuple(_label(int(g)) for g in (m.pair_geom1[p], m.pair_geom2[p]))
            if names not in allow_unsupported:
                raise mjcf.MJCFError("collision pair %s (geom types %d, %d) has no device collider; pass it in allow_unsupported to drop it knowingly"
                                     % (names, int(m.geom_type[int(m.pair_geom1[p])]), int(m.geom_type[int(m.pair_geom2[p])])))
            dropped_pairs.append(names)
    pair_order = [p for p in range(m.npair) if _ct(p) not in (CT_NONE, CT_CAP_ELL, CT_ELL_ELL)] + \
                 [p for p in range(m.npair) if _ct(p) in (CT_CAP_ELL, CT_ELL_ELL)]
    n_analytic = sum(1 for p in pair_order if _ct(p) not in (CT_CAP_ELL, CT_ELL_ELL))
    for p in pair_order:
        g1, g2 = int(m.pair_geom1[p]), int(m.pair_geom2[p])
        ct = _ct(p)
        dim = int(m.pair_dim[p])
        if dim not in (1, 3):
            raise mjcf.MJCFError("condim %d not supported" % dim)
        b1, b2 = int(m.geom_bodyid[g1]), int(m.geom_bodyid[g2])
        ds = sorted(set(chains.get(b1, ())) ^ set(chains.get(b2, ())))
        path_adr = len(PPATH)
        for d in ds:
            PPATH.append(d * 2 + (1 if moves(d, b2) else 0))
        maxpath = max(maxpath, len(ds))
        K, B, si = _kbimp(m.pair_solref[p], m.pair_solimp[p], h)
        tran = m.body_invweight0[b1, 0] + m.body_invweight0[b2, 0]
        crow = np.array([m.pair_margin[p], m.pair_gap[p], m.pair_friction[p, 0], m.pair_friction[p, 1], 0.0, K, B, *si])
        ckey = crow.tobytes()
        if ckey not in pcls_index:
            pcls_index[ckey] = len(PPAIR_d); PPAIR_d.append(crow)
        PPAIR.append([geom_ref(g1), geom_ref(g2), dim, path_adr, len(ds), ct, pcls_index[ckey], p])      # [7]: model pair index = rank in MuJoCo contact order
        PPAIR_tran.append(tran)
        pair_model_index.append(p)
    # ---- joint limits
    PLIM, PLIM_d = [], []
    for j in range(m.njnt):
        if m.jnt_limited[j] and m.jnt_type[j] in (mjcf.JNT_HINGE, mjcf.JNT_SLIDE):
            K, B, si = _kbimp(m.jnt_solref[j], m.jnt_solimp[j], h)
            d = int(m.jnt_dofadr[j])
            PLIM.append([d, int(m.jnt_qposadr[j])])
            PLIM_d.append([m.jnt_range[j, 0], m.jnt_range[j, 1], m.jnt_margin[j], m.dof_invweight0[d], K, B, *si, 0.0])
    # ---- joint equalities
    PEQ, PEQ_d = [], []
    eq_tree = 1
    for e in range(m.neq):
        if not m.eq_active0[e]:
            continue
        j1, j2 = int(m.eq_obj1id[e]), int(m.eq_obj2id[e])
        K, B, si = _kbimp(m.eq_solref[e], m.eq_solimp[e], h)
        q1, d1 = int(m.jnt_qposadr[j1]), int(m.jnt_dofadr[j1])
        q2, d2 = (int(m.jnt_qposadr[j2]), int(m.jnt_dofadr[j2])) if j2 >= 0 else (-1, -1)
        iw = m.dof_invweight0[d1] + (m.dof_invweight0[d2] if j2 >= 0 else 0.0)
        i12 = -1 if j2 < 0 else eidx.get((d1, d2), eidx.get((d2, d1), -2))
        if i12 == -2:
            eq_tree = 0
        PEQ.append([q1, d1, q2, d2, i12 if i12 >= 0 else -1, 0])
        PEQ_d.append([*m.eq_data[e], m.qpos0[q1], m.qpos0[q2] if j2 >= 0 else 0.0, iw, K, B, *si, 0.0])

    # The tree-sparse L'DL lists are dead weight in shared memory when the kernel can never take a sparse path: an equality that
    # couples dofs across the tree keeps every Newton Hessian dense (and nefc > 0), and the integrator uses the dense register
    # Cholesky for 8 <= nv <= 36 (myo_solver.cuh: phase_solve / phase_integrate).  Legs: 14 KB of the 74 KB hot blob.
    if len(PEQ) > 0 and not eq_tree and 8 <= m.nv <= 36:
        PLV_adr, PLV, PFE_adr, PFE, PFT_adr, PFT, PDS_adr, PDS = [0], [], [0], [], [0], [], [0], []
    dims = np.zeros(NPDIM, np.int32)
    dims[PD_NBD], dims[PD_NLEVEL], dims[PD_NPT], dims[PD_NSP], dims[PD_NWE] = nbd, nlevel, len(PPT_body), nsp, nwe
    dims[PD_NTA], dims[PD_NNZ], dims[PD_NTERM] = len(act_tendons), len(PNZ_dof), len(PTERM) // 3
    dims[PD_NLIM], dims[PD_NEQ], dims[PD_NPAIR], dims[PD_NGC] = len(PLIM), len(PEQ), len(PPAIR), len(PG_body)
    dims[PD_MAXPATH], dims[PD_MAXCHAIN], dims[PD_NSUB], dims[PD_NROW], dims[PD_NCOL] = maxpath, maxchain, len(PSUB), len(PROW_col), len(PCOL)
    dims[PD_NPIECE] = len(PT_piece)
    dims[PD_NWE_SPH_OUT], dims[PD_NWE_SPH_IN], dims[PD_NWE_CYL_OUT], dims[PD_NWE_CYL_IN] = counts
    dims[PD_NDEPTH], dims[PD_EQ_TREE], dims[PD_NPAIR_ANALYTIC] = ndepth, eq_tree, n_analytic
    # limit rows that can be active at once: both sides of a joint only when its range is narrower than twice the margin
    dims[PD_NLIMROW] = sum(2 if (r[1] - r[0]) < 2 * r[2] else 1 for r in PLIM_d)
    dims[PD_SPLIT_SP], dims[PD_SPLIT_WE], dims[PD_SPLIT_TA] = sp0[1], we0[1], split
    dims[PD_SPLIT_NZ] = sum(1 for t_ in PNZ_tendon if t_ < split)

    def ia(x, shape=None):
        a = np.asarray(x, dtype=np.int32)
        return a.reshape(shape) if shape else a

    prog = {
        "P_dims": dims,
        "PB_level_adr": ia(level_adr), "PB_parent": ia(PB_parent), "PB_jadr": ia(PB_jadr), "PB_jnum": ia(PB_jnum),
        "PB_model_id": ia(order), "PB_d": PB_d,
        "PD_body": ia(PD_body), "PD_lin": ia(PD_lin), "PDOF_d": PDOF_d,
        "PCH_adr": ia(PCH_adr), "PCH": ia(PCH), "PSUB_adr": ia(PSUB_adr), "PSUB": ia(PSUB),
        "PM_i": ia(PM_i), "PM_j": ia(PM_j), "PROW_adr": ia(PROW_adr), "PROW_col": ia(PROW_col), "PROW_idx": ia(PROW_idx),
        "PPT_body": ia(PPT_body), "PPT_xyz": np.array(PPT_xyz, dtype=np.float64).reshape(-1, 3),
        "PSP": ia(sp_list).reshape(-1, 2), "PWE": PWE, "PWE_d": PWE_d,
        "PT_const": np.array(T_const, dtype=np.float64), "PT_piece_adr": ia(PT_piece_adr), "PT_piece": ia(PT_piece),
        "PT_nz_adr": ia(PT_nz_adr), "PNZ_dof": ia(PNZ_dof), "PNZ_tendon": ia(PNZ_tendon), "PNZ_term_adr": ia(PNZ_term_adr),
        "PTERM": ia(PTERM), "PCOL_adr": ia(PCOL_adr), "PCOL": ia(PCOL),
        "PA_tendon": ia(PA_tendon), "PA_d": np.array(cls_rows, dtype=np.float64).reshape(-1, PA_STRIDE), "PA_cls": ia(PA_cls), "PAM_d": PAM_d,
        "PG_body": ia(PG_body), "PG_type": ia(PG_type), "PG_d": np.array(PG_d, dtype=np.float64).reshape(-1, PG_STRIDE),
        "PPAIR": ia(PPAIR).reshape(-1, PPAIR_ISTRIDE), "PPAIR_d": np.array(PPAIR_d, dtype=np.float64).reshape(-1, PPAIR_STRIDE),
        "PPAIR_tran": np.array(PPAIR_tran, dtype=np.float64),
        "PPATH": ia(PPATH),
        "PLIM": ia(PLIM).reshape(-1, 2), "PLIM_d": np.array(PLIM_d, dtype=np.float64).reshape(-1, PLIM_STRIDE),
        "PEQ": ia(PEQ).reshape(-1, 6),
        "PLV_adr": ia(PLV_adr), "PLV": ia(PLV), "PFE_adr": ia(PFE_adr), "PFE": ia(PFE), "PFT_adr": ia(PFT_adr), "PFT": ia(PFT),
        "PDS_adr": ia(PDS_adr), "PDS": ia(PDS), "PEQ_d": np.array(PEQ_d, dtype=np.float64).reshape(-1, PEQ_STRIDE),
    }
    info = dict(dyn_body_ids=order, act_tendons=act_tendons, pair_model_index=pair_model_index, dropped_pairs=dropped_pairs,
                geom_model_ids={v: k for k, v in gmap.items()})
    return prog, info


def _same_rigid(m, b0, b1):
    """True if the two bodies are rigidly attached (same weld body)."""
    return m.body_weldid[b0] == m.body_weldid[b1]


def _site_world(m, kin, s):
    b = m.site_bodyid[s]
    return kin["xpos"][b] + kin["xmat"][b] @ m.site_pos[s]
