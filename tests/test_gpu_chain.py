"""GPU parity: CUDA chaining (DP fill + backtrack with exact unstable-radix-sort tie order + compaction) vs the oracle."""
import numpy as np
import pytest
import oracle_lib as O
from test_oracle_vs_ref import make_anchors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import minimap2_b200 as mb
    c = mb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("cfg", [(5000, 5000, 500, 25, 5000, 3, 40, 0), (2000, 2000, 2000, 25, 50, 3, 100, 0), (200000, 2000, 200000, 25, 5000, 3, 40, 1)])
def test_chain_random(ctx, cfg):
    from minimap2_b200 import kernels as K
    mdx, mdy, bw, skip, iters, mincnt, minsc, is_cdna = cfg
    rng = np.random.default_rng(mdx + bw)
    arrs = [make_anchors(rng, n_chain=int(rng.integers(1, 6)), n_noise=int(rng.integers(0, 600))) for _ in range(150)]
    arrs.append(np.zeros((0, 2), dtype=np.uint64))  # empty read
    arrs.append(make_anchors(rng, n_chain=1, n_noise=0)[:2])  # fewer than min_cnt
    pg = float(np.float32(np.float32(0.8) * 0.01 * 15))
    got = K.chain_batch(ctx, arrs, mdx, mdy, bw, skip, iters, mincnt, minsc, pg, 0.0, is_cdna)
    for i, a in enumerate(arrs):
        if len(a) == 0:
            assert len(got[i][0]) == 0
            continue
        u, b = O.oracle_lchain_dp(a, mdx, mdy, bw, skip, iters, mincnt, minsc, pg, 0.0, is_cdna)
        assert len(u) == len(got[i][0]) and (u == got[i][0]).all(), i
        assert b.shape == got[i][1].shape and (b == got[i][1]).all(), i


@pytest.mark.parametrize("cfg", [(5000, 1000, 20000, 25, 100000, 3, 40), (800, 0, 500, 25, 100000, 3, 40), (5000, 1000, 2000, 5, 12, 2, 20)])
def test_chain_rmq_random(ctx, cfg):
    """mg_lchain_rmq kernel (AVL/RMQ trees in HBM arenas, one thread per read) vs the oracle restatement"""
    from minimap2_b200 import kernels as K
    md, mdi, bw, skip, cap, mincnt, minsc = cfg
    rng = np.random.default_rng(md + cap)
    arrs = [make_anchors(rng, n_chain=int(rng.integers(1, 5)), n_noise=int(rng.integers(0, 500))) for _ in range(80)]
    arrs.append(np.zeros((0, 2), dtype=np.uint64))
    pg = float(np.float32(np.float32(0.8) * 0.01 * 15))
    got = K.chain_rmq_batch(ctx, arrs, md, mdi, bw, skip, cap, mincnt, minsc, pg, 0.0)
    for i, a in enumerate(arrs):
        if len(a) == 0:
            assert len(got[i][0]) == 0
            continue
        u, b = O.oracle_lchain_rmq(a, md, mdi, bw, skip, cap, mincnt, minsc, pg, 0.0)
        assert len(u) == len(got[i][0]) and (u == got[i][0]).all(), i
        assert b.shape == got[i][1].shape and (b == got[i][1]).all(), i
