"""Random cases for the per-hit tail of the alignment driver (mm_append_cigar + mm_fix_cigar + mm_update_extra, align.c:105-334) and
runners for the three implementations the tests compare: the reference's own static functions (oracle/_ref/libminimap2_refalign.so),
the plain-C oracle (oracle/mm2o_extra.c) and the CUDA kernel K4 through the C-ABI (mmb_tail_batch_host). Test infrastructure."""
import ctypes as C
import os
import numpy as np
import oracle_lib as O

REFALIGN_SO = os.path.join(O.ORACLE_DIR, "_ref", "libminimap2_refalign.so")
COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def make_case(rng, low_complexity=True):
    """One hit: returns dict(read=forward read nt4, rev, qs, target=nt4 window incl. flanks, t0, pieces=[uint32 arrays], qspan, tspan)."""
    alpha = 2 if (low_complexity and rng.random() < 0.5) else 4  # two-letter stretches make most gaps shiftable
    ops = []
    n_ops = int(rng.integers(1, 40))
    for i in range(n_ops):
        r = rng.random()
        if r < 0.55 or i == 0 and rng.random() < 0.7: ops.append((0, int(rng.integers(1, 30))))
        elif r < 0.75: ops.append((1, int(rng.integers(1, 6))))
        elif r < 0.95: ops.append((2, int(rng.integers(1, 6))))
        elif r < 0.98: ops.append((int(rng.integers(0, 3)), 0))  # an empty operation (a piece boundary can leave one)
        else: ops.append((2, int(rng.integers(20, 90))))  # a long deletion: the logarithmic gap term
    ops.append((0, int(rng.integers(1, 12))))  # never an all-empty list (the reference would read past an empty CIGAR)
    tseq, qseq = [], []
    for op, ln in ops:
        if op == 0:
            seg = rng.integers(0, alpha, ln).astype(np.uint8)
            q = seg.copy()
            mm = rng.random(ln) < 0.08
            q[mm] = (q[mm] + 1 + rng.integers(0, 3, int(mm.sum()))) % 4
            tseq.append(seg); qseq.append(q)
        elif op == 1: qseq.append(rng.integers(0, alpha, ln).astype(np.uint8))
        else: tseq.append(rng.integers(0, alpha, ln).astype(np.uint8))
    t = np.concatenate(tseq) if tseq else np.zeros(0, np.uint8)
    q = np.concatenate(qseq) if qseq else np.zeros(0, np.uint8)
    for s in (t, q):  # a few ambiguous bases
        if len(s) and rng.random() < 0.3: s[rng.integers(0, len(s), int(rng.integers(1, 3)))] = 4
    # pieces: cut the operation list at random places, sometimes through an operation (mm_append_cigar merges it back)
    words = []
    cuts = set(int(x) for x in rng.integers(0, len(ops) + 1, int(rng.integers(0, 5))))
    pieces, cur = [], []
    for i, (op, ln) in enumerate(ops):
        if i in cuts and cur: pieces.append(cur); cur = []
        if ln > 1 and rng.random() < 0.15:
            a = int(rng.integers(1, ln)); cur.append(a << 4 | op); pieces.append(cur); cur = [(ln - a) << 4 | op]
        else: cur.append(ln << 4 | op)
    if cur: pieces.append(cur)
    if rng.random() < 0.2: pieces.insert(int(rng.integers(0, len(pieces) + 1)), [])  # an empty piece
    pieces = [np.array(p, dtype=np.uint32) for p in pieces]
    lf, rf = int(rng.integers(0, 20)), int(rng.integers(0, 20))
    sq = np.concatenate([rng.integers(0, 4, lf).astype(np.uint8), q, rng.integers(0, 4, rf).astype(np.uint8)])
    rev = int(rng.integers(0, 2))
    read = COMP[sq[::-1]] if rev else sq
    tl, tr = int(rng.integers(0, 20)), int(rng.integers(8, 30))
    target = np.concatenate([rng.integers(0, 4, tl).astype(np.uint8), t, rng.integers(0, 4, tr).astype(np.uint8)])
    return dict(read=read, rev=rev, qs=lf, qseq=q, tseq=t, target=target, t0=tl, pieces=pieces, qspan=len(q), tspan=len(t))


MAT = np.array([2, -4, -4, -4, -1, -4, 2, -4, -4, -1, -4, -4, 2, -4, -1, -4, -4, -4, 2, -1, -1, -1, -1, -1, -1], dtype=np.int8)  # map-ont


def _run_c(fn, case, q, e):
    pl = np.array([len(p) for p in case["pieces"]], dtype=np.uint32)
    ops = np.concatenate(case["pieces"] + [np.zeros(1, np.uint32)]).astype(np.uint32)
    qs_, ts_ = np.concatenate([case["qseq"], np.zeros(16, np.uint8)]), np.concatenate([case["tseq"], np.zeros(16, np.uint8)])
    qlen = len(case["read"])
    coor = np.array([qlen - case["qs"] - case["qspan"], qlen - case["qs"], 1000, 1000 + case["tspan"]] if case["rev"] else
                    [case["qs"], case["qs"] + case["qspan"], 1000, 1000 + case["tspan"]], dtype=np.int32)
    c0 = coor.copy()
    out = np.zeros(6, dtype=np.int32)
    cig = np.zeros(int(pl.sum()) + 4, dtype=np.uint32)
    fn(len(pl), pl.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_void_p), qs_.ctypes.data_as(C.c_void_p), ts_.ctypes.data_as(C.c_void_p),
       MAT.ctypes.data_as(C.c_void_p), C.c_int(q), C.c_int(e), C.c_int(case["rev"]), coor.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
       cig.ctypes.data_as(C.c_void_p))
    qshift = int(c0[1] - coor[1]) if case["rev"] else int(coor[0] - c0[0])
    return dict(n_cigar=int(out[0]), blen=int(out[1]), mlen=int(out[2]), n_ambi=int(out[3]), dp_max=int(out[4]), is_spliced=int(out[5]),
                qshift=qshift, tshift=int(coor[2] - c0[2]), cigar=cig[:out[0]].copy())


def run_oracle(case, q=4, e=2):
    L = O.oracle()
    L.mm2o_hit_tail.restype = None
    return _run_c(L.mm2o_hit_tail, case, q, e)


_refalign = None


def run_reference(case, q=4, e=2):
    global _refalign
    if _refalign is None:
        _refalign = C.CDLL(REFALIGN_SO)
        _refalign.refshim_hit_tail.restype = None
    return _run_c(_refalign.refshim_hit_tail, case, q, e)


class TailHit(C.Structure):  # mmb_tail_hit_t (include/mm_b200.h)
    _fields_ = [("q0", C.c_int64), ("t0", C.c_int64), ("qlen", C.c_int32), ("qs", C.c_int32), ("rev", C.c_int32), ("qspan", C.c_int32),
                ("tspan", C.c_int32), ("piece_first", C.c_int32), ("n_pieces", C.c_int32), ("pad", C.c_int32)]


class TailOut(C.Structure):  # mmb_tail_out_t
    _fields_ = [("n_cigar", C.c_int32), ("blen", C.c_int32), ("mlen", C.c_int32), ("n_ambi", C.c_int32), ("dp_max", C.c_int32),
                ("qshift", C.c_int32), ("tshift", C.c_int32), ("status", C.c_int32), ("is_spliced", C.c_int32), ("pad", C.c_int32 * 3)]


def run_device(ctx, lib, cases, q=4, e=2):
    """All cases in one K4 launch through the C-ABI (mmb_tail_batch_host). Returns a list of result dicts."""
    n = len(cases)
    hits = (TailHit * n)()
    reads, targets, plen, ops = [], [], [], []
    qo = to = po = 0
    cig_off = np.zeros(n + 1, dtype=np.int64)
    for i, c in enumerate(cases):
        h = hits[i]
        h.q0, h.t0, h.qlen, h.qs, h.rev, h.qspan, h.tspan = qo, to + c["t0"], len(c["read"]), c["qs"], c["rev"], c["qspan"], c["tspan"]
        h.piece_first, h.n_pieces = po, len(c["pieces"])
        reads.append(c["read"]); targets.append(c["target"])
        for p in c["pieces"]:
            plen.append(len(p)); ops.append(p)
        qo += len(c["read"]); to += len(c["target"]); po += len(c["pieces"])
        cig_off[i + 1] = cig_off[i] + sum(len(p) for p in c["pieces"])
    query = np.concatenate(reads + [np.zeros(1, np.uint8)]); target = np.concatenate(targets + [np.zeros(1, np.uint8)])
    plen = np.array(plen + [0], dtype=np.uint32); ops = np.concatenate(ops + [np.zeros(1, np.uint32)]).astype(np.uint32)
    out = (TailOut * n)()
    cig = np.zeros(int(cig_off[n]) + 4, dtype=np.uint32)
    lib.mmb_tail_batch_host.restype = C.c_int
    rc = lib.mmb_tail_batch_host(ctx.h, C.c_int(n), hits, C.c_int64(po), plen.ctypes.data_as(C.c_void_p), ops.ctypes.data_as(C.c_void_p),
                                 query.ctypes.data_as(C.c_void_p), C.c_int64(len(query)), target.ctypes.data_as(C.c_void_p), C.c_int64(len(target)),
                                 MAT.ctypes.data_as(C.c_void_p), C.c_int(q), C.c_int(e), cig_off.ctypes.data_as(C.c_void_p), out, cig.ctypes.data_as(C.c_void_p))
    assert rc == 0
    res = []
    for i in range(n):
        o = out[i]
        assert o.status == 0, (i, o.status)
        res.append(dict(n_cigar=o.n_cigar, blen=o.blen, mlen=o.mlen, n_ambi=o.n_ambi, dp_max=o.dp_max, is_spliced=o.is_spliced, qshift=o.qshift,
                        tshift=o.tshift, cigar=cig[cig_off[i]:cig_off[i] + o.n_cigar].copy()))
    return res


def same(a, b):
    return all(a[k] == b[k] for k in ("n_cigar", "blen", "mlen", "n_ambi", "dp_max", "is_spliced", "qshift", "tshift")) and np.array_equal(a["cigar"], b["cigar"])


def pack_results(res):
    """results -> flat arrays for an .npz fixture"""
    st = np.array([[r[k] for k in ("n_cigar", "blen", "mlen", "n_ambi", "dp_max", "is_spliced", "qshift", "tshift")] for r in res], dtype=np.int32)
    cg = np.concatenate([r["cigar"] for r in res] + [np.zeros(0, np.uint32)]).astype(np.uint32)
    return st, cg


def unpack_results(st, cg):
    res, o = [], 0
    for row in st:
        d = dict(zip(("n_cigar", "blen", "mlen", "n_ambi", "dp_max", "is_spliced", "qshift", "tshift"), (int(x) for x in row)))
        d["cigar"] = cg[o:o + d["n_cigar"]]; o += d["n_cigar"]
        res.append(d)
    return res
