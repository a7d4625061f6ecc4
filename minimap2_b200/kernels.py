"""Kernel-level host-buffer entry points (the C-ABI calls in include/mm_b200.h), used by the parity tests."""
import ctypes as C
import numpy as np
from ._lib import lib, KswJob, KswRes, KswScore


def make_score(mat, q, e, q2, e2):
    sc = KswScore()
    for i in range(25):
        sc.mat[i] = int(mat[i])
    sc.q, sc.e, sc.q2, sc.e2 = q, e, q2, e2
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
