"""CPU: the product's host-side hit logic (minimap2_b200/csrc/hits.cc -- chains -> hits, primary/secondary marking, secondary
selection, divergence estimate, sorting, SAM-primary marking, MAPQ) against the unmodified reference functions of hit.c /
esterr.c, struct for struct (every byte of mm_reg1_t, bit fields and the hash-seeded tie breakers included).
The product file is compiled as plain host C++ into a test shim (tests/hostshim); no GPU is involved."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import oracle_lib as O
from test_oracle_vs_ref import make_anchors

ROOT = O.ROOT
SHIM = os.path.join(ROOT, "tests", "hostshim", "_build", "libhostshim.so")
pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


class IdxSeq(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint64), ("len", C.c_uint32), ("is_alt", C.c_uint32)]


class Idx(C.Structure):  # mm_idx_t (minimap.h:88-100)
    _fields_ = [("b", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("flag", C.c_int32), ("n_seq", C.c_uint32),
                ("index", C.c_int32), ("n_alt", C.c_int32), ("seq", C.POINTER(IdxSeq)), ("S", C.POINTER(C.c_uint32)),
                ("B", C.c_void_p), ("I", C.c_void_p), ("spsc", C.c_void_p), ("J", C.c_void_p), ("km", C.c_void_p), ("h", C.c_void_p)]


REG_SIZE = 80


@pytest.fixture(scope="module")
def libs():
    os.makedirs(os.path.dirname(SHIM), exist_ok=True)
    src = [os.path.join(ROOT, "minimap2_b200", "csrc", "hits.cc"), os.path.join(ROOT, "tests", "hostshim", "hostshim.cc")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(s) > os.path.getmtime(SHIM) for s in src):
        inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "minimap2_b200", "csrc"), "-I/usr/local/cuda/include"]
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared"] + inc + src + ["-o", SHIM])
    H = C.CDLL(SHIM)
    R = O.ref()
    H.hs_gen_regs.restype = C.c_void_p
    R.mm_gen_regs.restype = C.c_void_p
    return H, R


def regs_bytes(ptr, n):
    return C.string_at(ptr, n * REG_SIZE) if n > 0 else b""


def make_case(rng, qlen=10000):
    a = make_anchors(rng, n_chain=int(rng.integers(1, 7)), n_noise=int(rng.integers(0, 300)), qlen=qlen)
    pg = float(np.float32(np.float32(0.8) * 0.01 * 15))
    u, b = O.ref_lchain_dp(a, 5000, 5000, 500, 25, 5000, 3, 40, pg, 0.0)
    return u, b


def mini_pos_of(rng, b, qlen):
    """forward query coordinates of the chain anchors plus some unrelated minimizers, as mm_collect_matches would list them"""
    xs = set()
    for x, y in b:
        q = int(np.int32(np.uint32(y & np.uint64(0xffffffff)))); span = int(y >> np.uint64(32) & np.uint64(0xff))
        xs.add(qlen - 1 - (q + 1 - span) if int(x >> np.uint64(63)) else q)
    for _ in range(int(rng.integers(0, 200))):
        xs.add(int(rng.integers(15, qlen)))
    return np.array([(15 << 32) | v for v in sorted(xs)], dtype=np.uint64)


@pytest.mark.parametrize("seed", range(6))
def test_hit_logic_matches_reference(libs, seed):
    H, R = libs
    rng = np.random.default_rng(700 + seed)
    names = [b"chr0", b"chr1", b"chr2"]
    seqs = (IdxSeq * 3)(*[IdxSeq(n, 0, 250000, 0) for n in names])
    mi = Idx(14, 10, 15, 0, 3, 0, 0, seqs, None, None, None, None, None, None, None)
    for it in range(60):
        qlen = 10000
        u, b = make_case(rng, qlen)
        if len(u) == 0:
            continue
        hash_ = int(rng.integers(0, 1 << 32))
        n = len(u)
        uh, ur = u.copy(), u.copy()
        bh, br = np.ascontiguousarray(b.copy()), np.ascontiguousarray(b.copy())
        ph = H.hs_gen_regs(C.c_uint32(hash_), qlen, n, uh.ctypes.data_as(C.c_void_p), bh.ctypes.data_as(C.c_void_p), 0)
        pr = R.mm_gen_regs(None, C.c_uint32(hash_), qlen, n, ur.ctypes.data_as(C.c_void_p), br.ctypes.data_as(C.c_void_p), 0)
        assert regs_bytes(ph, n) == regs_bytes(pr, n), ("gen_regs", it)
        assert (bh == br).all()
        # mm_set_parent + mm_select_sub as chain_post does (map.c:206-213), map-ont parameters
        for lib, p in ((H, ph), (R, pr)):
            if lib is H:
                lib.hs_set_parent(C.c_float(0.5), 0x7fffffff, n, C.c_void_p(p), 8, 0, C.c_float(0.15))
            else:
                lib.mm_set_parent(None, C.c_float(0.5), 0x7fffffff, n, C.c_void_p(p), 8, 0, C.c_float(0.15))
        assert regs_bytes(ph, n) == regs_bytes(pr, n), ("set_parent", it)
        nh, nr = C.c_int(n), C.c_int(n)
        H.hs_select_sub(C.c_float(0.8), 30, 5, 1, 4000, C.byref(nh), C.c_void_p(ph))
        R.mm_select_sub(None, C.c_float(0.8), 30, 5, 1, 4000, C.byref(nr), C.c_void_p(pr))
        assert nh.value == nr.value and regs_bytes(ph, nh.value) == regs_bytes(pr, nr.value), ("select_sub", it)
        n2 = nh.value
        # divergence estimate (esterr.c:30-64)
        mp = mini_pos_of(rng, b, qlen)
        H.hs_est_err(C.byref(mi), qlen, n2, C.c_void_p(ph), bh.ctypes.data_as(C.c_void_p), len(mp), mp.ctypes.data_as(C.c_void_p))
        R.mm_est_err(C.byref(mi), qlen, n2, C.c_void_p(pr), br.ctypes.data_as(C.c_void_p), len(mp), mp.ctypes.data_as(C.c_void_p))
        assert regs_bytes(ph, n2) == regs_bytes(pr, n2), ("est_err", it)
        a_ = H.hs_filter_strand_retained(n2, C.c_void_p(ph)); b_ = R.mm_filter_strand_retained(n2, C.c_void_p(pr))
        assert a_ == b_ and regs_bytes(ph, a_) == regs_bytes(pr, b_), ("filter_strand_retained", it)
        n3 = a_
        # no base-level alignment: mm_set_mapq2 on chaining scores (map.c:338-343), then sorting and SAM primary flags
        rep_len = int(rng.integers(0, 3000))
        H.hs_set_mapq(n3, C.c_void_p(ph), 40, 2, rep_len, 0, 0)
        R.mm_set_mapq2(None, n3, C.c_void_p(pr), 40, 2, rep_len, 0, 0)
        assert regs_bytes(ph, n3) == regs_bytes(pr, n3), ("set_mapq", it)
        nh, nr = C.c_int(n3), C.c_int(n3)
        H.hs_hit_sort(C.byref(nh), C.c_void_p(ph), C.c_float(0.15))
        R.mm_hit_sort(None, C.byref(nr), C.c_void_p(pr), C.c_float(0.15))
        assert nh.value == nr.value and regs_bytes(ph, nh.value) == regs_bytes(pr, nr.value), ("hit_sort", it)
        assert H.hs_set_sam_pri(nh.value, C.c_void_p(ph)) == R.mm_set_sam_pri(nr.value, C.c_void_p(pr))
        assert regs_bytes(ph, nh.value) == regs_bytes(pr, nr.value), ("set_sam_pri", it)
        H.hs_free(C.c_void_p(ph)); O.ref().refshim_free(C.c_void_p(pr))


def test_sdust_matches_reference(libs):
    """symmetric DUST (sdust.c) restated in csrc/hits.cc vs the reference's sdust(): random, low-complexity, tandem-repeat and
    N-interrupted sequences, several thresholds and window sizes"""
    H, R = libs[0], libs[1]
    R.sdust.restype = C.POINTER(C.c_uint64)
    R.sdust.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    H.hs_sdust.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(9)
    n_masked = 0
    for it in range(120):
        parts = []
        for _ in range(int(rng.integers(1, 8))):
            kind = int(rng.integers(0, 5))
            if kind == 0:
                parts.append(bytes(rng.choice(list(b"ACGT"), int(rng.integers(5, 400)))))
            elif kind == 1:
                parts.append(bytes([int(rng.choice(list(b"ACGT")))]) * int(rng.integers(3, 120)))
            elif kind == 2:
                unit = bytes(rng.choice(list(b"ACGT"), int(rng.integers(2, 7)))); parts.append(unit * int(rng.integers(2, 40)))
            elif kind == 3:
                parts.append(b"N" * int(rng.integers(1, 5)))
            else:
                parts.append(bytes(rng.choice(list(b"AAAT"), int(rng.integers(10, 200)))).lower())
        seq = b"".join(parts)
        T = int(rng.choice([10, 20, 20, 30, 50])); W = int(rng.choice([16, 32, 64, 64, 100]))
        n = C.c_int(0)
        p = R.sdust(None, seq, len(seq), T, W, C.byref(n))
        ref = [int(p[i]) for i in range(n.value)]
        R.refshim_free(p)
        out = (C.c_uint64 * 4096)()
        m = H.hs_sdust(seq, len(seq), T, W, out, 4096)
        assert m == n.value and [int(out[i]) for i in range(m)] == ref, (it, T, W, seq[:80])
        n_masked += m
    assert n_masked > 300
