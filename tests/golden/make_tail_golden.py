"""Generates tests/golden/vectors_tail.npz: outputs of the REFERENCE's own per-hit tail (mm_append_cigar + mm_fix_cigar + mm_update_extra,
static in align.c, reached through oracle/_ref/libminimap2_refalign.so) on the seeded random cases of tests/tail_cases.py. Run here (the
reference exists only in this container); the fixture travels to the GPU box: python tests/golden/make_tail_golden.py"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import tail_cases as T

SEED, N = 20260923, 600
rng = np.random.default_rng(SEED)
cases = [T.make_case(rng) for _ in range(N)]
st, cg = T.pack_results([T.run_reference(c) for c in cases])
np.savez_compressed(os.path.join(HERE, "vectors_tail.npz"), seed=np.int64(SEED), n=np.int64(N), stats=st, cigars=cg)
print("vectors_tail.npz: %d cases, %d CIGAR operations" % (N, len(cg)))
