"""CPU: the chaining kernels (chain_fill_kernel: warp-per-read DP with scan-based skip counter; chain_bt_kernel: warp-per-read
backtrack/compaction; chain_rescue_kernel: mg_lchain_rmq with the AVL/RMQ tree) -- unmodified CUDA sources under the SIMT emulator --
against the oracle on random anchor sets with ties, several chains per read, empty reads."""
import ctypes as C
import os
import sys
import numpy as np
import pytest
import oracle_lib as O
from test_oracle_vs_ref import make_anchors

sys.path.insert(0, os.path.join(O.ROOT, "tests", "cuda_emu"))
sys.path.insert(0, O.ROOT)


@pytest.fixture(scope="module")
def emu():
    import build_emu
    from minimap2_b200._lib import ChainPar
    L = C.CDLL(build_emu.build("mmb_emu_all", build_emu.ALL, extra=()))
    L.mmb_ctx_create.restype = C.c_void_p
    ctx = C.c_void_p(L.mmb_ctx_create(0))
    return L, ctx, ChainPar


def run(emu, fn, arrs, par):
    L, ctx, ChainPar = emu
    n = len(arrs)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, a in enumerate(arrs):
        off[i + 1] = off[i] + len(a)
    tot = int(off[-1])
    cat = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.uint64).reshape(-1, 2) for a in arrs]))
    n_u = np.zeros(n, dtype=np.int32); n_v = np.zeros(n, dtype=np.int32)
    u = np.zeros(tot + 1, dtype=np.uint64); ao = np.zeros((tot + 1, 2), dtype=np.uint64)
    getattr(L, fn)(ctx, C.byref(par), n, C.c_void_p(cat.ctypes.data), C.c_void_p(off.ctypes.data), C.c_void_p(n_u.ctypes.data), C.c_void_p(n_v.ctypes.data),
                   C.c_void_p(u.ctypes.data), C.c_void_p(ao.ctypes.data))
    return [(u[int(off[i]):int(off[i]) + n_u[i]].copy(), ao[int(off[i]):int(off[i]) + n_v[i]].copy()) for i in range(n)]


@pytest.mark.parametrize("cfg", [(5000, 5000, 500, 25, 5000, 3, 40, 0), (200000, 2000, 200000, 25, 60, 3, 40, 1)])
def test_emulated_chain_dp_matches_oracle(emu, cfg):
    mdx, mdy, bw, skip, iters, mincnt, minsc, is_cdna = cfg
    rng = np.random.default_rng(mdx + bw)
    arrs = [make_anchors(rng, n_chain=int(rng.integers(1, 4)), n_noise=int(rng.integers(0, 120)), qlen=4000) for _ in range(6)]
    arrs.append(np.zeros((0, 2), dtype=np.uint64))
    arrs.append(make_anchors(rng, n_chain=1, n_noise=0)[:2])
    pg = float(np.float32(np.float32(0.8) * 0.01 * 15))
    par = emu[2](mdx, mdy, bw, skip, iters, mincnt, minsc, pg, 0.0, is_cdna, 1, 0, 0, 0)
    got = run(emu, "mmb_chain_batch_host", arrs, par)
    for i, a in enumerate(arrs):
        if len(a) == 0:
            assert len(got[i][0]) == 0
            continue
        u, b = O.oracle_lchain_dp(a, mdx, mdy, bw, skip, iters, mincnt, minsc, pg, 0.0, is_cdna)
        assert len(u) == len(got[i][0]) and (u == got[i][0]).all(), i
        assert b.shape == got[i][1].shape and (b == got[i][1]).all(), i


def test_emulated_chain_rmq_matches_oracle(emu):
    rng = np.random.default_rng(99)
    arrs = [make_anchors(rng, n_chain=int(rng.integers(1, 4)), n_noise=int(rng.integers(0, 120)), qlen=4000) for _ in range(5)]
    pg = float(np.float32(np.float32(0.8) * 0.01 * 15))
    for (md, mdi, bw, skip, cap, mincnt, minsc) in [(5000, 1000, 20000, 25, 100000, 3, 40), (5000, 1000, 2000, 5, 12, 2, 20)]:
        par = emu[2](md, md, bw, skip, 5000, mincnt, minsc, pg, 0.0, 0, 1, 1, mdi, cap)
        got = run(emu, "mmb_chain_rmq_batch_host", arrs, par)
        for i, a in enumerate(arrs):
            u, b = O.oracle_lchain_rmq(a, md, mdi, bw, skip, cap, mincnt, minsc, pg, 0.0)
            assert len(u) == len(got[i][0]) and (u == got[i][0]).all(), (md, cap, i)
            assert b.shape == got[i][1].shape and (b == got[i][1]).all(), (md, cap, i)
