"""GPU parity for K4 (finalize.cu): the per-hit tail of the alignment driver -- CIGAR stitching (mm_append_cigar), mm_fix_cigar and
mm_update_extra (align.c:105-334) -- through the C-ABI entry mmb_tail_batch_host against the oracle (oracle/mm2o_extra.c) and against
the committed outputs of the reference's own functions (tests/golden/vectors_tail.npz). Bit-exact: CIGAR words, blen, mlen, n_ambi,
dp_max, the coordinate shifts of a dropped leading gap."""
import os
import numpy as np
import pytest
import tail_cases as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors_tail.npz")


@pytest.fixture(scope="module")
def ctx():
    import minimap2_b200 as mb
    c = mb.Context(0)
    yield c
    c.close()


def test_tail_matches_oracle(ctx):
    import minimap2_b200 as mb
    rng = np.random.default_rng(77)
    cases = [T.make_case(rng) for _ in range(3000)]
    got = T.run_device(ctx, mb.lib(), cases)
    bad = [i for i, c in enumerate(cases) if not T.same(got[i], T.run_oracle(c))]
    assert not bad, bad[:10]


def test_tail_other_gap_costs(ctx):  # the logarithmic gap term with other q / e
    import minimap2_b200 as mb
    rng = np.random.default_rng(78)
    cases = [T.make_case(rng) for _ in range(500)]
    for q, e in ((6, 2), (5, 4), (16, 1)):
        got = T.run_device(ctx, mb.lib(), cases, q, e)
        bad = [i for i, c in enumerate(cases) if not T.same(got[i], T.run_oracle(c, q, e))]
        assert not bad, (q, e, bad[:10])


def test_tail_matches_reference_vectors(ctx):
    import minimap2_b200 as mb
    z = np.load(GOLD)
    rng = np.random.default_rng(int(z["seed"]))
    cases = [T.make_case(rng) for _ in range(int(z["n"]))]
    exp = T.unpack_results(z["stats"], z["cigars"])
    got = T.run_device(ctx, mb.lib(), cases)
    bad = [i for i in range(len(cases)) if not T.same(got[i], exp[i])]
    assert not bad, bad[:10]


def test_tail_long_hit(ctx):  # one hit of read scale: ~40 pieces, thousands of operations
    import minimap2_b200 as mb
    rng = np.random.default_rng(79)
    big = []
    for _ in range(8):
        parts = [T.make_case(rng, low_complexity=False) for _ in range(40)]
        q = np.concatenate([p["qseq"] for p in parts]); t = np.concatenate([p["tseq"] for p in parts])
        pieces = [np.concatenate(p["pieces"]) if p["pieces"] else np.zeros(0, np.uint32) for p in parts]
        big.append(dict(read=q, rev=0, qs=0, qseq=q, tseq=t, target=np.concatenate([t, np.zeros(16, np.uint8)]), t0=0, pieces=pieces, qspan=len(q), tspan=len(t)))
    got = T.run_device(ctx, mb.lib(), big)
    for g, c in zip(got, big):
        assert T.same(g, T.run_oracle(c))
