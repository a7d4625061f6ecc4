"""CPU: kernel K4 (finalize.cu, unmodified CUDA source under the SIMT emulator) against the oracle's per-hit tail (oracle/mm2o_extra.c)."""
import ctypes as C
import os
import sys
import numpy as np
import pytest
import oracle_lib as O
import tail_cases as T

sys.path.insert(0, os.path.join(O.ROOT, "tests", "cuda_emu"))
sys.path.insert(0, O.ROOT)


class _Ctx:
    def __init__(self, h):
        self.h = h


def test_emulated_tail_matches_oracle():
    import build_emu
    L = C.CDLL(build_emu.build("mmb_emu_all", build_emu.ALL, extra=()))
    L.mmb_ctx_create.restype = C.c_void_p
    ctx = _Ctx(C.c_void_p(L.mmb_ctx_create(0)))
    rng = np.random.default_rng(11)
    cases = [T.make_case(rng) for _ in range(300)]
    got = T.run_device(ctx, L, cases)
    bad = [i for i, c in enumerate(cases) if not T.same(got[i], T.run_oracle(c))]
    assert not bad, bad[:10]
