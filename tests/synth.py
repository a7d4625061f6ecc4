"""Deterministic synthetic genomes/reads for tests and bench (SURVEY 8d shapes, small sizes here)."""
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}


def random_genome(n, seed, n_contigs=1, repeat_frac=0.0):
    rng = np.random.default_rng(seed)
    g = ALPHA[rng.integers(0, 4, n)]
    if repeat_frac > 0:  # copy segments around to create repeats (exercises mid_occ / chaining)
        n_rep = int(n * repeat_frac / 3000)
        for _ in range(n_rep):
            L = int(rng.integers(1000, 5000)); s = int(rng.integers(0, n - L)); d = int(rng.integers(0, n - L))
            g[d:d + L] = g[s:s + L]
    bounds = np.linspace(0, n, n_contigs + 1).astype(int)
    return [g[bounds[i]:bounds[i + 1]].copy() for i in range(n_contigs)]


def mutate_ascii(seq, rng, err, sub=0.4, ins=0.25):
    """vectorised ONT-like errors on an ASCII uint8 array"""
    n = len(seq)
    r = rng.random(n); kind = rng.random(n)
    is_err = r < err
    is_sub = is_err & (kind < sub)
    is_ins = is_err & (kind >= sub) & (kind < sub + ins)
    is_del = is_err & (kind >= sub + ins)
    codes = np.searchsorted(ALPHA, seq)  # ACGT -> 0..3
    newc = (codes + 1 + rng.integers(0, 3, n)) % 4
    out = seq.copy()
    out[is_sub] = ALPHA[newc[is_sub]]
    reps = np.ones(n, dtype=np.int64)
    reps[is_del] = 0
    reps[is_ins] = 2
    res = np.repeat(out, reps)
    # inserted base = second copy at insertion sites: randomise it
    pos = np.cumsum(reps) - 1
    ins_pos = pos[is_ins]
    res[ins_pos] = ALPHA[rng.integers(0, 4, len(ins_pos))]
    return res


def revcomp(a):
    lut = np.zeros(256, dtype=np.uint8)
    for k, v in COMP.items():
        lut[k] = v
    return lut[a[::-1]]


def make_reads(contigs, n_reads, read_len, err, seed, chimeric_frac=0.0):
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(n_reads):
        c = contigs[int(rng.integers(0, len(contigs)))]
        L = min(read_len, len(c))
        s = int(rng.integers(0, len(c) - L + 1))
        seg = c[s:s + L]
        if chimeric_frac > 0 and rng.random() < chimeric_frac:
            c2 = contigs[int(rng.integers(0, len(contigs)))]
            s2 = int(rng.integers(0, len(c2) - L // 2 + 1))
            seg = np.concatenate([seg[:L // 2], c2[s2:s2 + L // 2]])
        rd = mutate_ascii(seg, rng, err)
        if rng.random() < 0.5:
            rd = revcomp(rd)
        reads.append(rd)
    return reads


def write_fasta(path, names, seqs):
    with open(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + n.encode() + b"\n")
            f.write(bytes(s) + b"\n")
