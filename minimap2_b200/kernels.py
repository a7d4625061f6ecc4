"""Kernel-level host-buffer entry points (the C-ABI calls in include/mm_b200.h), used by the parity tests."""
import ctypes as C
import numpy as np
from ._lib import lib, KswJob, KswRes, KswScore, ChainPar


def make_score(mat, q, e, q2, e2, noncan=0, junc_bonus=0, junc_pen=0, zd_skip=0):
    sc = KswScore()
    for i in range(25):
        sc.mat[i] = int(mat[i])
    sc.q, sc.e, sc.q2, sc.e2 = q, e, q2, e2
    sc.noncan, sc.junc_bonus, sc.junc_pen = noncan, junc_bonus, junc_pen
    sc.zd_skip = zd_skip
    return sc


def ksw_batch(ctx, score, pairs, params):
    """pairs: list of (q, t) uint8 nt4 arrays; params: list of dict(w, zdrop, end_bonus, flag).
    Returns a list of dicts shaped like the oracle's (tests/oracle_lib.ez_dict)."""
    n = len(pairs)
    if n == 0:
        return []
    qcat = np.concatenate([np.asarray(p[0], dtype=np.uint8) for p in pairs])
    tcat = np.concatenate([np.asarray(p[1], dtype=np.uint8) for p in pairs])
    jobs = (KswJob * n)()
    qo = to = 0
    tot = 0
    for i, ((q, t), pr) in enumerate(zip(pairs, params)):
        j = jobs[i]
        j.q_start, j.t_start, j.q_step, j.t_step = qo, to, 1, 1
        j.qlen, j.tlen = len(q), len(t)
        j.w, j.zdrop, j.end_bonus, j.flag = pr["w"], pr["zdrop"], pr["end_bonus"], pr["flag"]
        qo += len(q); to += len(t); tot += len(q) + len(t) + 2
    res = (KswRes * n)()
    cig = np.zeros(tot, dtype=np.uint32)
    used = lib().mmb_ksw_batch_host(ctx.h, C.byref(score), n, jobs, qcat.ctypes.data, len(qcat), tcat.ctypes.data, len(tcat),
                                    res, cig.ctypes.data, len(cig))
    assert used >= 0, used
    out = []
    for i in range(n):
        r = res[i]
        out.append(dict(max=r.max, zdropped=r.zdropped, max_q=r.max_q, max_t=r.max_t, mqe=r.mqe, mqe_t=r.mqe_t, mte=r.mte,
                        mte_q=r.mte_q, score=r.score, n_cigar=r.n_cigar, reach_end=r.reach_end,
                        cigar=[int(x) for x in cig[r.cigar_off:r.cigar_off + r.n_cigar]]))
    return out


def sketch_batch(ctx, seqs, w, k, is_hpc=0, rid0=0):
    """seqs: list of bytes. Returns a list of (n_i, 2) uint64 arrays (x, y) like mm_sketch's mm128_t output."""
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    cat = b"".join(seqs)
    buf = np.frombuffer(cat, dtype=np.uint8)
    n_out = np.zeros(n, dtype=np.int64)
    cap = max(int(off[-1]), 1)
    out = np.zeros((cap, 2), dtype=np.uint64)
    tot = lib().mmb_sketch_batch_host(ctx.h, n, buf.ctypes.data, off.ctypes.data, w, k, is_hpc, rid0, out.ctypes.data, cap, n_out.ctypes.data)
    assert tot <= cap
    res, o = [], 0
    for i in range(n):
        res.append(out[o:o + n_out[i]].copy()); o += int(n_out[i])
    assert o == tot
    return res


def chain_batch(ctx, anchor_arrays, max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, pen_gap, pen_skip=0.0, is_cdna=0, n_seg=1):
    """anchor_arrays: list of (n_i,2) uint64 arrays sorted by x. Returns list of (u, a) like mg_lchain_dp."""
    n = len(anchor_arrays)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, a in enumerate(anchor_arrays):
        off[i + 1] = off[i] + len(a)
    tot = int(off[-1])
    cat = np.concatenate([np.asarray(a, dtype=np.uint64).reshape(-1, 2) for a in anchor_arrays]) if tot else np.zeros((0, 2), dtype=np.uint64)
    cat = np.ascontiguousarray(cat)
    par = ChainPar(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, pen_gap, pen_skip, is_cdna, n_seg, 0, 0, 0)
    n_u = np.zeros(n, dtype=np.int32); n_v = np.zeros(n, dtype=np.int32)
    u = np.zeros(tot + 1, dtype=np.uint64); ao = np.zeros((tot + 1, 2), dtype=np.uint64)
    lib().mmb_chain_batch_host(ctx.h, C.byref(par), n, cat.ctypes.data, off.ctypes.data, n_u.ctypes.data, n_v.ctypes.data, u.ctypes.data, ao.ctypes.data)
    out = []
    for i in range(n):
        o = int(off[i])
        out.append((u[o:o + n_u[i]].copy(), ao[o:o + n_v[i]].copy()))
    return out


def chain_rmq_batch(ctx, anchor_arrays, max_dist, max_dist_inner, bw, max_skip, cap, min_cnt, min_sc, pen_gap, pen_skip=0.0):
    """mg_lchain_rmq on every read (anchors sorted by x). Returns list of (u, a)."""
    n = len(anchor_arrays)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, a in enumerate(anchor_arrays):
        off[i + 1] = off[i] + len(a)
    tot = int(off[-1])
    cat = np.concatenate([np.asarray(a, dtype=np.uint64).reshape(-1, 2) for a in anchor_arrays]) if tot else np.zeros((0, 2), dtype=np.uint64)
    cat = np.ascontiguousarray(cat)
    par = ChainPar(max_dist, max_dist, bw, max_skip, 5000, min_cnt, min_sc, pen_gap, pen_skip, 0, 1, 1, max_dist_inner, cap)
    n_u = np.zeros(n, dtype=np.int32); n_v = np.zeros(n, dtype=np.int32)
    u = np.zeros(tot + 1, dtype=np.uint64); ao = np.zeros((tot + 1, 2), dtype=np.uint64)
    lib().mmb_chain_rmq_batch_host(ctx.h, C.byref(par), n, cat.ctypes.data, off.ctypes.data, n_u.ctypes.data, n_v.ctypes.data, u.ctypes.data, ao.ctypes.data)
    out = []
    for i in range(n):
        o = int(off[i])
        out.append((u[o:o + n_u[i]].copy(), ao[o:o + n_v[i]].copy()))
    return out
