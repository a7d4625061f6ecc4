"""CPU: the minimap.h boundary itself under the SIMT emulator -- mm_idx_str / mm_set_opt / mm_mapopt_update / mm_tbuf_init / mm_map
of the emulated product library next to the same calls on the unmodified reference library, on identical inputs. Every
mm_reg1_t and mm_extra_t field (and the tbuf's rep_len) must be equal: this is example.c's / mappy's call sequence
(example.c:33-59, cmappy.h:74-109) run against both implementations."""
import ctypes as C
import os
import sys
import numpy as np
import pytest
import oracle_lib as O
import synth
from test_hostlogic_vs_ref import Idx, REG_SIZE
from test_aligndriver_vs_ref import extra_of

sys.path.insert(0, os.path.join(O.ROOT, "tests", "cuda_emu"))
sys.path.insert(0, O.ROOT)
pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def test_mm_map_matches_reference_library():
    import build_emu
    from minimap2_b200 import api
    E = C.CDLL(build_emu.build("mmb_emu_all", build_emu.ALL, extra=()))
    R = O.ref()
    contigs = synth.random_genome(60_000, 61, n_contigs=2, repeat_frac=0.1)
    reads = [bytes(r) for r in synth.make_reads(contigs, 3, 1500, 0.08, 161, chimeric_frac=0.0)]
    reads.append(b"ACGTACGTAC")   # shorter than k: no hit (map.c:243)
    names = [b"chrA", b"chrB"]
    outs = []
    for lib in (E, R):
        lib.mm_idx_str.restype = C.POINTER(Idx)
        lib.mm_map.restype = C.c_void_p
        lib.mm_tbuf_init.restype = C.c_void_p
        lib.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(api.IdxOpt), C.POINTER(api.MapOpt)]
        seqs = (C.c_char_p * 2)(*[bytes(c) for c in contigs]); nms = (C.c_char_p * 2)(*names)
        io, mo = api.IdxOpt(), api.MapOpt()
        lib.mm_set_opt(None, C.byref(io), C.byref(mo)); lib.mm_set_opt(b"map-ont", C.byref(io), C.byref(mo))
        mo.flag |= api.MM_F_CIGAR
        mi = lib.mm_idx_str(io.w, io.k, 0, io.bucket_bits, 2, seqs, nms)
        lib.mm_mapopt_update(C.byref(mo), mi)
        tb = C.c_void_p(lib.mm_tbuf_init())
        res = []
        for qi, rd in enumerate(reads):
            n = C.c_int(-1)
            regs = lib.mm_map(mi, len(rd), rd, C.byref(n), tb, C.byref(mo), b"read%d" % qi)
            recs = []
            for i in range(n.value):
                recs.append((C.string_at(regs + i * REG_SIZE, 72), extra_of(regs, i)))
            res.append((n.value, bool(regs), recs))
        outs.append((mo.mid_occ, res))
        lib.mm_tbuf_destroy(tb)
    assert outs[0][0] == outs[1][0]                    # mm_mapopt_update derived the same mid_occ from the two indexes
    for qi, (a, b) in enumerate(zip(outs[0][1], outs[1][1])):
        assert a[0] == b[0], (qi, a[0], b[0])   # (for n == 0 the reference hands back a zero-length malloc, this library NULL: both are free()-able)
        for i, (x, y) in enumerate(zip(a[2], b[2])):
            assert x == y, (qi, i, np.frombuffer(x[0], dtype=np.int32), np.frombuffer(y[0], dtype=np.int32))
    assert outs[0][1][0][0] >= 1 and outs[0][1][3][0] == 0
