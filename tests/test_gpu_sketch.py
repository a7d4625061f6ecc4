"""GPU parity: chunk-parallel CUDA minimizer sketch vs the oracle restatement of mm_sketch (bit-exact, order included)."""
import numpy as np
import pytest
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import minimap2_b200 as mb
    c = mb.Context(0)
    yield c
    c.close()


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(list(alphabet), n).astype(np.uint8))


@pytest.mark.parametrize("w,k,hpc", [(10, 15, 0), (5, 15, 0), (19, 19, 0), (10, 14, 0), (11, 21, 0), (10, 19, 1), (3, 4, 0), (50, 28, 0), (200, 6, 0)])
def test_sketch_random(ctx, w, k, hpc):
    from minimap2_b200 import kernels as K
    rng = np.random.default_rng(w * 1000 + k)
    seqs = []
    for it in range(80):
        n = int(rng.integers(1, 6000))
        s = rand_seq(rng, n, rng.choice([b"ACGT", b"ACGTN", b"AT", b"ACGTacgtNn", b"AC"]))
        if rng.random() < 0.3:
            s = s[: n // 2] + b"AT" * 60 + b"A" * 50 + b"N" * int(rng.integers(0, 40)) + s[n // 2:]
        seqs.append(s)
    seqs.append(rand_seq(rng, 40000))                       # many chunks
    seqs.append(rand_seq(rng, 30000, b"ACGTN"))             # N everywhere: every chunk falls back
    seqs.append(b"ACGT" * 3000)                             # period-4 repeat (symmetric k-mers when k even)
    seqs.append(b"AT" * 5000)
    seqs.append(b"A")
    got = K.sketch_batch(ctx, seqs, w, k, hpc, rid0=7)
    for i, s in enumerate(seqs):
        exp = O.oracle_sketch(s, w, k, rid=7 + i, is_hpc=hpc)
        assert got[i].shape == exp.shape and (got[i] == exp).all(), (i, len(s), got[i].shape, exp.shape)


def test_sketch_read_batch_shape(ctx):
    """the batch shape the mapper uses: thousands of 10 kb reads, k15 w10"""
    from minimap2_b200 import kernels as K
    rng = np.random.default_rng(9)
    seqs = [rand_seq(rng, 10000) for _ in range(300)]
    got = K.sketch_batch(ctx, seqs, 10, 15)
    for i in range(0, 300, 7):
        exp = O.oracle_sketch(seqs[i], 10, 15, rid=i)
        assert (got[i] == exp).all()
