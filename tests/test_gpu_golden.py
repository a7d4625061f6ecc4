"""GPU parity against the committed golden fixtures (tests/golden): the CLI's PAF/SAM output vs the reference binary's
recorded output, and the CUDA kernels (through the C-ABI, host buffers) vs the recorded outputs of the reference functions.
Nothing here needs oracle/_ref or /root/reference."""
import os
import subprocess
import numpy as np
import pytest
import oracle_lib as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(O.ROOT, "tests", "golden")
MINE = os.path.join(O.ROOT, "minimap2_b200", "minimap2-b200")


def load_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CASES


# the spliced cases live in tests/test_gpu_zz_splice.py (newest path, run last)
@pytest.mark.parametrize("name", sorted(k for k in load_cases().keys() if not k.startswith("splice")))
def test_cli_output_matches_recorded_reference(name):
    args = load_cases()[name]
    p = subprocess.run([MINE, "-t", "8"] + args, cwd=os.path.join(GOLD, "data"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    got = [l for l in p.stdout.decode().splitlines() if not l.startswith("@PG")]
    exp = open(os.path.join(GOLD, "expected", name + ".txt")).read().splitlines()
    assert len(got) == len(exp), (len(got), len(exp))
    for i, (a, b) in enumerate(zip(exp, got)):
        assert a == b, "line %d differs:\nref: %s\ngot: %s" % (i, a[:600], b[:600])


@pytest.fixture(scope="module")
def ctx():
    import minimap2_b200 as mb
    c = mb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def V():
    return np.load(os.path.join(GOLD, "vectors.npz"))


def test_sketch_kernel_matches_golden(ctx, V):
    from minimap2_b200 import kernels as K
    nc, ns = V["sk_n"]
    for ci in range(nc):
        for si in range(ns):
            w, k, hpc, rid = [int(x) for x in V["sk%d_%d_par" % (ci, si)]]
            got = K.sketch_batch(ctx, [V["sk%d_%d_seq" % (ci, si)].tobytes()], w, k, hpc, rid0=rid)[0]
            exp = V["sk%d_%d_out" % (ci, si)]
            assert got.shape == exp.shape and (got == exp).all(), (ci, si)


def test_chain_kernel_matches_golden(ctx, V):
    from minimap2_b200 import kernels as K
    n = int(V["ch_n"][0])
    arrs = [V["ch%d_a" % i] for i in range(n)]
    par = [int(x) for x in V["ch0_par"]]
    pg, ps = [float(x) for x in V["ch0_pen"]]
    got = K.chain_batch(ctx, arrs, *par, pg, ps)
    for i in range(n):
        assert (got[i][0] == V["ch%d_u" % i]).all() and got[i][1].shape == V["ch%d_b" % i].shape and (got[i][1] == V["ch%d_b" % i]).all(), i


def test_chain_rmq_kernel_matches_golden(ctx, V):
    """mg_lchain_rmq (the long-join rescue chainer) against the reference function's recorded output"""
    from minimap2_b200 import kernels as K
    n = int(V["ch_n"][0])
    arrs = [V["ch%d_a" % i] for i in range(n)]
    par = [int(x) for x in V["rq0_par"]]
    pg, ps = [float(x) for x in V["ch0_pen"]]
    got = K.chain_rmq_batch(ctx, arrs, *par, pg, ps)
    for i in range(n):
        assert len(got[i][0]) == len(V["rq%d_u" % i]) and (got[i][0] == V["rq%d_u" % i]).all(), i
        assert got[i][1].shape == V["rq%d_b" % i].shape and (got[i][1] == V["rq%d_b" % i]).all(), i


KEYS = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "reach_end"]


def test_ksw_kernel_matches_golden(ctx, V):
    from minimap2_b200 import kernels as K
    n = int(V["kw_n"][0])
    pairs = [(V["kw%d_q" % i], V["kw%d_t" % i]) for i in range(n)]
    params = []
    for i in range(n):
        w, zdrop, end_bonus, flag = [int(x) for x in V["kw%d_par" % i]]
        params.append(dict(w=w, zdrop=zdrop, end_bonus=end_bonus, flag=flag))
    got = K.ksw_batch(ctx, K.make_score(V["kw_mat"], 4, 2, 24, 1), pairs, params)
    for i in range(n):
        exp = V["kw%d_res" % i]
        for k, e in zip(KEYS, exp):
            assert got[i][k] == int(e), (i, k, got[i][k], int(e), params[i])
        assert got[i]["cigar"] == [int(x) for x in V["kw%d_cig" % i]], i


def test_ksw_ll_kernel_matches_golden(ctx, V):
    """ksw_ll_i16 (ksw2_ll_sse.c:85; the inversion probe of align.c:930-987): jobs flagged MMB_JOB_LL carry gap open/extend in w/zdrop"""
    from minimap2_b200 import kernels as K
    MMB_JOB_LL = 0x20000
    n = int(V["ll_n"][0])
    pairs = [(V["ll%d_q" % i], V["ll%d_t" % i]) for i in range(n)]
    params = [dict(w=4, zdrop=2, end_bonus=0, flag=MMB_JOB_LL) for _ in range(n)]
    got = K.ksw_batch(ctx, K.make_score(V["kw_mat"], 4, 2, 24, 1), pairs, params)
    for i in range(n):
        sc, qe, te = [int(x) for x in V["ll%d_res" % i]]
        assert (got[i]["score"], got[i]["max_q"], got[i]["max_t"]) == (sc, qe, te), (i, got[i], (sc, qe, te))
