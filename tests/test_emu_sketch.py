"""CPU: the sketch kernels (sketch_tile_kernel over 2-bit-packed bases -- the production path for odd k; sketch_kernel chunk replay --
even k / HPC / wide windows), unmodified CUDA sources under the SIMT emulator (the TMA bulk copy is replaced by a plain staging loop
there), against the oracle restatement of mm_sketch: values and order."""
import ctypes as C
import os
import sys
import numpy as np
import pytest
import oracle_lib as O

sys.path.insert(0, os.path.join(O.ROOT, "tests", "cuda_emu"))
sys.path.insert(0, O.ROOT)


@pytest.fixture(scope="module")
def emu():
    import build_emu
    L = C.CDLL(build_emu.build("mmb_emu_all", build_emu.ALL, extra=()))
    L.mmb_ctx_create.restype = C.c_void_p
    L.mmb_sketch_batch_host.restype = C.c_int64
    return L, C.c_void_p(L.mmb_ctx_create(0))


def sketch(emu, seqs, w, k, hpc=0, rid0=0):
    L, ctx = emu
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    buf = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8)
    n_out = np.zeros(n, dtype=np.int64)
    cap = max(int(off[-1]), 1)
    out = np.zeros((cap, 2), dtype=np.uint64)
    tot = L.mmb_sketch_batch_host(ctx, n, C.c_void_p(buf.ctypes.data), C.c_void_p(off.ctypes.data), w, k, hpc, rid0, C.c_void_p(out.ctypes.data),
                                  C.c_int64(cap), C.c_void_p(n_out.ctypes.data))
    res, o = [], 0
    for i in range(n):
        res.append(out[o:o + n_out[i]].copy()); o += int(n_out[i])
    assert o == tot
    return res


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(list(alphabet), n).astype(np.uint8))


@pytest.mark.parametrize("w,k", [(10, 15), (5, 15), (19, 19), (11, 21), (1, 15), (36, 27), (49, 15), (10, 14)])
def test_emulated_sketch_matches_oracle(emu, w, k):
    rng = np.random.default_rng(w * 100 + k)
    seqs = []
    for it in range(24):
        n = int(rng.integers(1, 3000))
        s = rand_seq(rng, n, [b"ACGT", b"ACGTN", b"AT", b"ACGTacgtNn", b"AC", b"A"][int(rng.integers(0, 6))])
        if rng.random() < 0.4:  # low complexity: ties inside the window, duplicates, runs of N
            s = s[: n // 2] + b"AT" * 40 + b"A" * 70 + b"N" * int(rng.integers(0, 40)) + b"ACG" * 30 + s[n // 2:]
        seqs.append(s)
    seqs.append(rand_seq(rng, 2048 * 2 + 5))       # tile boundaries
    seqs.append(rand_seq(rng, 2048))
    seqs.append(rand_seq(rng, 2047, b"ACGTN"))
    seqs.append(b"A" * 4200)
    seqs.append(b"")
    seqs.append(b"ACGT" * 700)
    seqs.append(b"N" * 100 + rand_seq(rng, 300) + b"N")
    got = sketch(emu, seqs, w, k, rid0=3)
    for i, s in enumerate(seqs):
        exp = O.oracle_sketch(s, w, k, rid=3 + i) if len(s) else np.zeros((0, 2), dtype=np.uint64)
        assert got[i].shape == exp.shape and (got[i] == exp).all(), (i, len(s), got[i].shape, exp.shape)


def test_emulated_sketch_output_arena_overflow_is_recovered(emu):
    """the output arena is sized from the expected minimizer density; when a batch yields more (MM_B200_SKETCH_CAP shrinks the arena
    here), the kernel reports the exact total and is launched again with enough room"""
    rng = np.random.default_rng(5)
    seqs = [rand_seq(rng, int(rng.integers(3000, 5000))) for _ in range(3)]
    os.environ["MM_B200_SKETCH_CAP"] = "100"
    try:
        got = sketch(emu, seqs, 3, 5)
    finally:
        del os.environ["MM_B200_SKETCH_CAP"]
    for i, s in enumerate(seqs):
        exp = O.oracle_sketch(s, 3, 5, rid=i)
        assert got[i].shape == exp.shape and (got[i] == exp).all()
