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
    for _ in range(12):  # hits spanning several 32-operation chunks and more than 32 pieces: the chunk carries of every stage
        parts = [T.make_case(rng, low_complexity=bool(rng.integers(0, 2))) for _ in range(int(rng.integers(3, 14)))]
        q = np.concatenate([p["qseq"] for p in parts]); t = np.concatenate([p["tseq"] for p in parts])
        pieces = [x for p in parts for x in p["pieces"]]
        cases.append(dict(read=q, rev=0, qs=0, qseq=q, tseq=t, target=np.concatenate([t, np.zeros(16, np.uint8)]), t0=0, pieces=pieces, qspan=len(q), tspan=len(t)))
    got = T.run_device(ctx, L, cases)
    bad = [i for i, c in enumerate(cases) if not T.same(got[i], T.run_oracle(c))]
    assert not bad, bad[:10]
