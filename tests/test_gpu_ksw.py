"""GPU parity: the CUDA ksw2 extd2 kernel (through the C-ABI, host buffers) vs the oracle restatement, bit-exact."""
import numpy as np
import pytest
import oracle_lib as O

pytestmark = pytest.mark.gpu

EXT, RIGHT, REVC, APPROX, GENERIC = 0x40, 0x02, 0x80, 0x08, 0x04


@pytest.fixture(scope="module")
def ctx():
    import minimap2_b200 as mb
    c = mb.Context(0)
    yield c
    c.close()


def rand_pair(rng, qlen, err):
    t = rng.integers(0, 4, qlen + rng.integers(0, 30)).astype(np.uint8)
    q = O.mutate(t, rng, err=err)
    if len(q) == 0:
        q = np.array([0], dtype=np.uint8)
    return q, t


def compare(ctx, pairs, params, scoring=(2, 4, 4, 2, 24, 1), ts=0):
    from minimap2_b200 import kernels as K
    a, b, q, e, q2, e2 = scoring
    mat = O.simple_mat(a, b, 1, ts)
    got = K.ksw_batch(ctx, K.make_score(mat, q, e, q2, e2), pairs, params)
    for i, ((qq, tt), pr) in enumerate(zip(pairs, params)):
        exp = O.oracle_extd2(qq, tt, mat, q, e, q2, e2, pr["w"], pr["zdrop"], pr["end_bonus"], pr["flag"])
        if pr["flag"] & APPROX:
            exp = dict(exp)  # approx mode: ez->max family is untouched by the reference; identical by construction
        assert got[i] == exp, (i, len(qq), len(tt), pr, {k: (got[i][k], exp[k]) for k in exp if got[i][k] != exp[k]})


@pytest.mark.parametrize("seed", range(4))
def test_extd2_random_small(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    pairs, params = [], []
    for it in range(400):
        qlen = int(rng.integers(1, 450))
        q, t = rand_pair(rng, qlen, err=float(rng.choice([0.0, 0.05, 0.15, 0.4])))
        if rng.random() < 0.2:
            q[rng.integers(0, len(q))] = 4
            t[rng.integers(0, len(t))] = 4
        if rng.random() < 0.15:
            q = np.concatenate([q, rng.integers(0, 4, 300).astype(np.uint8)])
            t = np.concatenate([t, rng.integers(0, 4, 300).astype(np.uint8)])
        pairs.append((q, t))
        params.append(dict(w=int(rng.choice([-1, 5, 17, 40, 100, 751, 30001])), zdrop=int(rng.choice([-1, 50, 200, 400])),
                           end_bonus=int(rng.choice([-1, 0, 10])),
                           flag=int(rng.choice([0, APPROX, EXT, EXT | RIGHT | REVC, RIGHT, EXT | RIGHT, APPROX | RIGHT]))))
    compare(ctx, pairs, params)


def test_extd2_map_ont_shape(ctx):
    """the shape the mapper produces: ~230x230 gap fills w=30001 approx, plus end extensions w=751 exact"""
    rng = np.random.default_rng(77)
    pairs, params = [], []
    for it in range(600):
        t = rng.integers(0, 4, int(rng.integers(150, 560))).astype(np.uint8)
        q = O.mutate(t, rng, err=0.1)
        pairs.append((q, t))
        if it % 10 == 0:
            params.append(dict(w=751, zdrop=400, end_bonus=-1, flag=EXT | (RIGHT | REVC if it % 20 == 0 else 0)))
        else:
            params.append(dict(w=30001, zdrop=400, end_bonus=-1, flag=APPROX if it % 3 else 0))
    compare(ctx, pairs, params)


def test_extd2_long_band_limited(ctx):
    rng = np.random.default_rng(11)
    pairs, params = [], []
    for it in range(6):
        t = rng.integers(0, 4, int(rng.integers(1500, 5000))).astype(np.uint8)
        q = O.mutate(t, rng, err=0.12)
        for flag in (EXT, EXT | RIGHT | REVC, 0, APPROX):
            for w in (751, 100, 33):
                pairs.append((q, t)); params.append(dict(w=w, zdrop=400, end_bonus=-1, flag=flag))
    compare(ctx, pairs, params)


def test_extd2_other_scoring(ctx):
    rng = np.random.default_rng(5)
    for scoring, ts in [((1, 4, 6, 2, 26, 1), 0), ((1, 19, 39, 3, 81, 1), 0), ((2, 6, 10, 2, 50, 1), 4)]:
        pairs, params = [], []
        for it in range(100):
            q, t = rand_pair(rng, int(rng.integers(1, 300)), 0.1)
            pairs.append((q, t))
            params.append(dict(w=int(rng.choice([-1, 20, 200, 30001])), zdrop=200, end_bonus=-1,
                               flag=int(rng.choice([0, APPROX, EXT, EXT | RIGHT | REVC])) | (GENERIC if ts else 0)))
        compare(ctx, pairs, params, scoring, ts)


def test_extd2_edge_cases(ctx):
    one = np.array([1], dtype=np.uint8)
    pairs = [(one, one), (one, np.array([2, 1, 3], dtype=np.uint8)), (np.array([0, 1, 2, 3] * 8, dtype=np.uint8), one),
             (np.full(40, 4, dtype=np.uint8), np.full(33, 4, dtype=np.uint8))]
    params = [dict(w=-1, zdrop=400, end_bonus=-1, flag=f) for f in (0, EXT, APPROX, 0)]
    compare(ctx, pairs, params)


def test_zdrop_scan_and_skip(ctx):
    """the packed kernel's in-kernel mm_test_zdrop scan (align.c:61-89) on the device: exact against a plain restatement, and the
    scan-skip bound (mmb_ksw_score_t::zd_skip) only ever answers "no drop" when the true drop is within the threshold"""
    import minimap2_b200 as mb
    from minimap2_b200._lib import KswJob, KswRes, KswScore
    import test_emu_ksw as E
    dev = (mb.lib(), ctx.h, KswJob, KswRes, KswScore)
    mat = O.simple_mat(2, 4, 1)
    E.check_zdrop_scan_skip(lambda *a: E.run_jobs(dev, *a), mat)
    rng = np.random.default_rng(3)
    pairs = []
    for it in range(200):
        t = rng.integers(0, 4, int(rng.integers(30, 500))).astype(np.uint8)
        pairs.append((O.mutate(t, rng, err=float(rng.choice([0.05, 0.15, 0.3]))), t))
    pairs = [(q if len(q) else np.array([0], dtype=np.uint8), t) for q, t in pairs]
    params = [dict(w=30001, zdrop=400, end_bonus=-1, flag=E.APPROX | E.JOB_ZDROP)] * len(pairs)
    for (qq, tt), (g, zd) in zip(pairs, E.run_jobs(dev, mat, 4, 2, 24, 1, pairs, params)):
        exp = O.oracle_extd2(qq, tt, mat, 4, 2, 24, 1, 30001, 400, -1, E.APPROX)
        assert g == exp and zd == E.zdrop_scan(qq, tt, mat, exp["cigar"], 4, 2)
