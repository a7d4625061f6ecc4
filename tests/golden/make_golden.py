#!/usr/bin/env python3
"""Regenerates the committed golden fixtures from the UNMODIFIED reference built under oracle/_ref (make -C oracle ref).

    python tests/golden/make_golden.py

Two kinds of fixtures are written next to this script:

* expected/<case>.txt -- stdout of oracle/_ref/minimap2 (reference main.c over its own library) for the command lines in
  CASES, run on the FASTA files under data/ (copied test inputs of the reference: test/MT-*.fa, t-inv/q-inv, x3s-*, and a
  small synthetic ONT-like set written by this script). SAM @PG lines are dropped (they carry the program path).
* vectors.npz -- kernel-level known-answer vectors taken from the reference functions through oracle/ref_shim.c:
  mm_sketch (sketch.c:77), mg_lchain_dp (lchain.c:148), mg_lchain_rmq (lchain.c:251), ksw_extd2_sse (ksw2_extd2_sse.c:27),
  ksw_ll_i16 (ksw2_ll_sse.c:85) and radix_sort_128x (misc.c / ksort.h:98) on seeded inputs.

The tests never need /root/reference or oracle/_ref to check against these files: tests/test_golden.py pins the plain-C
oracle (oracle/mm2o_*.c) to vectors.npz on the CPU, and tests/test_gpu_golden.py compares the CUDA path with both.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
import synth  # noqa: E402

DATA = os.path.join(HERE, "data")
EXPECTED = os.path.join(HERE, "expected")

# name -> minimap2 argument list (paths relative to data/)
CASES = {
    "mt_paf_cigar": ["-c", "MT-human.fa", "MT-orang.fa"],
    "mt_sam": ["-a", "MT-human.fa", "MT-orang.fa"],
    "mt_paf_nocigar": ["MT-human.fa", "MT-orang.fa"],
    "mt_paf_cs_md": ["-c", "--cs", "--MD", "MT-human.fa", "MT-orang.fa"],
    "inv_paf_cigar": ["-c", "t-inv.fa", "q-inv.fa"],
    "x3s_paf_cigar": ["-c", "x3s-ref.fa", "x3s-qry.fa"],
    "t2_paf_cigar": ["-c", "t2.fa", "q2.fa"],
    "ont_paf_cs": ["-x", "map-ont", "-c", "--cs", "synth-ref.fa", "synth-ont.fa"],
    "ont_sam_md": ["-x", "map-ont", "-a", "--MD", "synth-ref.fa", "synth-ont.fa"],
    "ont_paf_nocigar_allchains": ["-x", "map-ont", "-P", "synth-ref.fa", "synth-ont.fa"],
    "hifi_paf_cigar": ["-x", "map-hifi", "-c", "synth-ref.fa", "synth-hifi.fa"],
    "mt_paf_single_affine": ["-c", "-O4", "-E2", "MT-human.fa", "MT-orang.fa"],          # q == q2, e == e2: ksw_extz2 in the reference
    "ont_paf_single_affine": ["-x", "map-ont", "-c", "-O6", "-E2", "synth-ref.fa", "synth-ont.fa"],
    "splice_paf_cs": ["-x", "splice", "-c", "--cs", "synth-gene.fa", "synth-cdna.fa"],          # ksw_exts2 path: introns, ts:A, both transcript strands
    "splice_sam_fwd": ["-x", "splice", "-uf", "-a", "--MD", "synth-gene.fa", "synth-cdna.fa"],   # forward transcript strand only
}


def write_synthetic_inputs():
    contigs = synth.random_genome(120_000, 11, n_contigs=2, repeat_frac=0.15)
    synth.write_fasta(os.path.join(DATA, "synth-ref.fa"), ["chrA", "chrB"], contigs)
    reads = synth.make_reads(contigs, 40, 2500, 0.10, 111, chimeric_frac=0.1)
    synth.write_fasta(os.path.join(DATA, "synth-ont.fa"), ["ont%d" % i for i in range(len(reads))], reads)
    reads = synth.make_reads(contigs, 12, 6000, 0.01, 112, chimeric_frac=0.0)
    synth.write_fasta(os.path.join(DATA, "synth-hifi.fa"), ["hifi%d" % i for i in range(len(reads))], reads)


def write_spliced_inputs():
    """a 60 kb 'gene region' with planted GT..AG / CT..AC signals and cDNA reads made of 2-4 exons (3 % errors), both strands"""
    rng = np.random.default_rng(3)
    contigs = synth.random_genome(60_000, 31, n_contigs=1, repeat_frac=0.0)
    g = np.frombuffer(bytes(contigs[0]), dtype=np.uint8).copy()
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    reads = []
    for i in range(10):
        pos = int(rng.integers(1000, len(g) - 12000)); exons = []; rev = i % 2 == 1
        for k in range(int(rng.integers(2, 5))):
            el = int(rng.integers(90, 300)); exons.append((pos, pos + el)); il = int(rng.integers(150, 1500))
            if i % 3 != 2:
                d, a = (b"GT", b"AG") if not rev else (b"CT", b"AC")
                g[pos + el:pos + el + 2] = list(d); g[pos + el + il - 2:pos + el + il] = list(a)
            pos += el + il
        tr = np.concatenate([g[s:e] for s, e in exons])
        if rev:
            tr = comp[tr[::-1]]
        reads.append(synth.mutate_ascii(tr, rng, 0.03))
    synth.write_fasta(os.path.join(DATA, "synth-gene.fa"), ["chr0"], [g.tobytes()])
    synth.write_fasta(os.path.join(DATA, "synth-cdna.fa"), ["tr%d" % i for i in range(len(reads))], reads)


def run_cases():
    os.makedirs(EXPECTED, exist_ok=True)
    for name, args in CASES.items():
        p = subprocess.run([O.REF_BIN, "-t", "2"] + args, cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
        lines = [l for l in p.stdout.decode().splitlines() if not l.startswith("@PG")]
        with open(os.path.join(EXPECTED, name + ".txt"), "w") as f:
            f.write("\n".join(lines) + ("\n" if lines else ""))
        print("%-28s %4d lines" % (name, len(lines)))


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(list(alphabet), n).astype(np.uint8))


def anchors_for_chain(rng, n_true, n_noise, qlen, tlen, strand_mix=True):
    """A plausible anchor set: a co-linear run with indel drift plus random noise, sorted like radix_sort_128x leaves it."""
    a = []
    q = rng.integers(20, 60); t = rng.integers(1000, tlen // 2)
    for _ in range(n_true):
        q += int(rng.integers(8, 40)); t += int(rng.integers(8, 40))
        if rng.random() < 0.02:
            t += int(rng.integers(100, 3000))   # a long gap for the rescue path
        if q >= qlen:
            break
        a.append((0, t, q))
    for _ in range(n_noise):
        a.append((int(rng.integers(0, 2)) if strand_mix else 0, int(rng.integers(0, tlen)), int(rng.integers(15, qlen))))
    out = np.zeros((len(a), 2), dtype=np.uint64)
    for i, (rev, t, q) in enumerate(a):
        out[i, 0] = (np.uint64(rev) << np.uint64(63)) | (np.uint64(0) << np.uint64(32)) | np.uint64(t)
        out[i, 1] = (np.uint64(15) << np.uint64(32)) | np.uint64(q)
    return O.ref_sort128(out)


def make_vectors():
    rng = np.random.default_rng(20240607)
    v = {}
    # ---- mm_sketch ----
    sk = [(10, 15, 0), (5, 15, 0), (19, 19, 0), (10, 19, 1), (3, 4, 0), (11, 21, 0)]
    for ci, (w, k, hpc) in enumerate(sk):
        for si in range(4):
            n = int(rng.integers(30, 3000))
            s = rand_seq(rng, n, [b"ACGT", b"ACGTN", b"AT", b"ACGTacgtNn"][si])
            if si == 1:
                s = s[: n // 2] + b"A" * 40 + b"AT" * 30 + s[n // 2:]
            v["sk%d_%d_seq" % (ci, si)] = np.frombuffer(s, dtype=np.uint8)
            v["sk%d_%d_par" % (ci, si)] = np.array([w, k, hpc, 3 + si], dtype=np.int32)
            v["sk%d_%d_out" % (ci, si)] = O.ref_sketch(s, w, k, rid=3 + si, is_hpc=hpc)
    v["sk_n"] = np.array([len(sk), 4], dtype=np.int32)
    # ---- radix_sort_128x (tie order is contract) ----
    for i, n in enumerate([10, 64, 65, 300, 5000]):
        a = np.zeros((n, 2), dtype=np.uint64)
        a[:, 0] = rng.integers(0, 50 if i % 2 else 1 << 40, n).astype(np.uint64) | (rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63))
        a[:, 1] = np.arange(n, dtype=np.uint64)
        v["rs%d_in" % i] = a
        v["rs%d_out" % i] = O.ref_sort128(a)
    v["rs_n"] = np.array([5], dtype=np.int32)
    # ---- mg_lchain_dp / mg_lchain_rmq ----
    nch = 6
    for i in range(nch):
        a = anchors_for_chain(rng, int(rng.integers(50, 400)), int(rng.integers(0, 600)), 8000, 200000)
        par = [5000, 5000, 500, 25, 5000, 3, 40]
        pen_gap, pen_skip = np.float32(0.8 * 0.01 * 15), np.float32(0.0)
        u, b = O.ref_lchain_dp(a, *par, float(pen_gap), float(pen_skip))
        v["ch%d_a" % i] = a
        v["ch%d_par" % i] = np.array(par, dtype=np.int32)
        v["ch%d_pen" % i] = np.array([pen_gap, pen_skip], dtype=np.float32)
        v["ch%d_u" % i] = u
        v["ch%d_b" % i] = b
        rpar = [5000, 1000, 20000, 25, 100000, 3, 40]
        u, b = O.ref_lchain_rmq(a, *rpar, float(pen_gap), float(pen_skip))
        v["rq%d_par" % i] = np.array(rpar, dtype=np.int32)
        v["rq%d_u" % i] = u
        v["rq%d_b" % i] = b
    v["ch_n"] = np.array([nch], dtype=np.int32)
    # ---- ksw_extd2_sse ----
    mat = O.simple_mat(2, 4, 1)
    cases = []
    KSW_EZ_SCORE_ONLY, KSW_EZ_RIGHT, KSW_EZ_APPROX_MAX, KSW_EZ_APPROX_DROP, KSW_EZ_EXTZ_ONLY, KSW_EZ_REV_CIGAR = 0x01, 0x02, 0x08, 0x10, 0x40, 0x80
    flags = [0, KSW_EZ_APPROX_MAX, KSW_EZ_EXTZ_ONLY, KSW_EZ_EXTZ_ONLY | KSW_EZ_RIGHT | KSW_EZ_REV_CIGAR, KSW_EZ_APPROX_MAX | KSW_EZ_APPROX_DROP,
             KSW_EZ_SCORE_ONLY, KSW_EZ_RIGHT]
    for i in range(42):
        tl = int(rng.integers(1, 400))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = O.mutate(t, rng, err=float(rng.choice([0.02, 0.1, 0.25])))
        if len(q) == 0:
            q = np.array([0], dtype=np.uint8)
        if i % 7 == 3:
            q[rng.integers(0, len(q))] = 4
        fl = flags[i % len(flags)]
        w = int(rng.choice([-1, 5, 30, 500]))
        zdrop = int(rng.choice([-1, 100, 400]))
        end_bonus = int(rng.choice([-1, 0, 10]))
        r = O.ref_extd2(q, t, mat, 4, 2, 24, 1, w, zdrop, end_bonus, fl)
        v["kw%d_q" % i] = q; v["kw%d_t" % i] = t
        v["kw%d_par" % i] = np.array([w, zdrop, end_bonus, fl], dtype=np.int32)
        v["kw%d_res" % i] = np.array([r["max"], r["zdropped"], r["max_q"], r["max_t"], r["mqe"], r["mqe_t"], r["mte"], r["mte_q"], r["score"], r["reach_end"]], dtype=np.int64)
        v["kw%d_cig" % i] = np.array(r["cigar"], dtype=np.uint32)
    v["kw_n"] = np.array([42], dtype=np.int32)
    v["kw_mat"] = np.asarray(mat, dtype=np.int8)
    # ---- ksw_ll_i16 ----
    for i in range(12):
        tl = int(rng.integers(5, 300))
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = O.mutate(t[int(rng.integers(0, tl // 2)):], rng, err=0.1)
        if len(q) == 0:
            q = np.array([1], dtype=np.uint8)
        sc, qe, te = O.ref_ll_i16(q, t, mat, 4, 2)
        v["ll%d_q" % i] = q; v["ll%d_t" % i] = t
        v["ll%d_res" % i] = np.array([sc, qe, te], dtype=np.int32)
    v["ll_n"] = np.array([12], dtype=np.int32)
    # ---- ksw_exts2_sse (spliced alignment; checker for the next hot-path row) ----
    from test_oracle_vs_ref import _spliced_pair
    smat = O.simple_mat(1, 2, 1)
    sflags = [0x100, 0x200 | 0x400, 0x100 | 0x08, 0x100 | 0x40, 0x200 | 0x40 | 0x02 | 0x80, 0x100 | 0x800, 0x100 | 0x02, 0x200 | 0x01]
    for i in range(24):
        q, t = _spliced_pair(rng, int(rng.integers(1, 5)), float(rng.choice([0.0, 0.03, 0.1])))
        fl = sflags[i % len(sflags)]
        zdrop = int(rng.choice([-1, 200])); end_bonus = int(rng.choice([-1, 10]))
        r = O.ref_exts2(q, t, smat, 2, 1, 32, 9, zdrop, end_bonus, 9, 5, fl)
        v["sp%d_q" % i] = q; v["sp%d_t" % i] = t
        v["sp%d_par" % i] = np.array([zdrop, end_bonus, fl], dtype=np.int32)
        v["sp%d_res" % i] = np.array([r["max"], r["zdropped"], r["max_q"], r["max_t"], r["mqe"], r["mqe_t"], r["mte"], r["mte_q"], r["score"], r["reach_end"]], dtype=np.int64)
        v["sp%d_cig" % i] = np.array(r["cigar"], dtype=np.uint32)
    v["sp_n"] = np.array([24], dtype=np.int32)
    v["sp_mat"] = np.asarray(smat, dtype=np.int8)
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **v)
    print("vectors.npz: %d arrays" % len(v))


if __name__ == "__main__":
    if not (O.have_ref() and os.path.exists(O.REF_BIN)):
        sys.exit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference is mounted")
    write_synthetic_inputs()
    write_spliced_inputs()
    run_cases()
    make_vectors()
