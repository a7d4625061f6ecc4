"""Differential fuzzing of the emulated CLI (tests/cuda_emu/_build/minimap2-emu: the product sources under the SIMT emulator) against the
reference binary oracle/_ref/minimap2: random small genomes (repeats, tandem blocks), reads with deletions / inversions / chimeric
joins on either strand, and random option sets (presets, output formats, -P -X --for-only --rev-only --qstrand -T -f a,b -e -k -w -A -B
-O -E -z -r -N -p -s -m -n -K -g ...). Every run must reproduce the reference output line for line. Development tool, not part of
the suite (a case takes seconds to minutes under emulation): python tests/cuda_emu/fuzz_cli.py [--splice] FIRST_SEED LAST_SEED"""
import sys, os, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth, oracle_lib as O
from concurrent.futures import ThreadPoolExecutor
EMU = os.path.join(ROOT, "tests", "cuda_emu", "_build", "minimap2-emu")  # built by tests/test_emu_e2e.py's fixture
env = dict(os.environ, MM_B200_GROUPS="1")
comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")

def one(seed):
    try:
        return one_(seed)
    except subprocess.TimeoutExpired:
        return seed, True, "TIMEOUT(emulation too slow)", 0, ""

def one_(seed):
    rng = np.random.default_rng(seed)
    d = tempfile.mkdtemp(prefix="fz%d_" % seed)
    glen = int(rng.integers(4000, 14000)); nc = int(rng.integers(1, 3))
    contigs = synth.random_genome(glen, seed, n_contigs=nc, repeat_frac=float(rng.choice([0, 0, 0.3])))
    gs = [np.frombuffer(bytes(c), dtype=np.uint8).copy() for c in contigs]
    for g in gs:  # some low-complexity and tandem copies
        if rng.random() < 0.5 and len(g) > 1500:
            st = int(rng.integers(0, len(g) - 600)); g[st:st + 60] = np.resize(g[st:st + int(rng.integers(2, 12))], 60)
    reads = []
    for i in range(int(rng.integers(2, 5))):
        g = gs[int(rng.integers(0, nc))]; L = int(rng.integers(600, min(3000, len(g) - 10))); st = int(rng.integers(0, len(g) - L))
        x = g[st:st + L]
        if rng.random() < 0.3 and L > 800:  # deletion / inversion / chimera
            k = int(rng.integers(0, 3)); m = L // 2
            if k == 0: x = np.concatenate([x[:m - 100], x[m + 100:]])
            elif k == 1: x = np.concatenate([x[:m - 150], comp[x[m - 150:m + 150][::-1]], x[m + 150:]])
            else:
                g2 = gs[int(rng.integers(0, nc))]; s2 = int(rng.integers(0, len(g2) - 400)); x = np.concatenate([x[:m], g2[s2:s2 + 400]])
        if rng.random() < 0.5: x = comp[x[::-1]]
        reads.append(synth.mutate_ascii(x, rng, float(rng.choice([0.0, 0.02, 0.05, 0.1]))))
    rf, qf = d + "/ref.fa", d + "/q.fa"
    synth.write_fasta(rf, ["c%d" % i for i in range(nc)], [g.tobytes() for g in gs]); synth.write_fasta(qf, ["r%d" % i for i in range(len(reads))], reads)
    args = []
    preset = str(rng.choice(["", "", "map-ont", "map-ont", "map-hifi", "asm20", "asm10", "map-pb", "lr:hq", "ava-ont"]))
    if preset: args += ["-x", preset]
    out = str(rng.choice(["-c", "-a", "", "-c --cs", "-c --MD --cs=long", "-a --eqx", "-c --ds"]))
    args += out.split()
    for opt, p in (("-P", .15), ("--for-only", .1), ("--rev-only", .1), ("-X", .1), ("--no-long-join", .15), ("-Y", .2), ("--secondary=no", .15), ("--paf-no-hit", .2)):
        if rng.random() < p: args.append(opt)
    if rng.random() < .3: args += ["-z", str(int(rng.choice([30, 60, 100]))) + "," + str(int(rng.choice([20, 40])))]
    if rng.random() < .3: args += ["-r", str(int(rng.choice([50, 200]))) + "," + str(int(rng.choice([500, 3000])))]
    if rng.random() < .3: args += ["-f", str(rng.choice(["2", "3,40", "0.01"]))]
    if rng.random() < .2: args += ["-e", str(int(rng.choice([0, 100])))]
    if rng.random() < .3: args += ["-k", str(int(rng.choice([11, 13, 17]))), "-w", str(int(rng.choice([3, 5, 8])))]
    if rng.random() < .25: args += ["-A", "1", "-B", str(int(rng.choice([2, 4]))), "-O", str(rng.choice(["3,12", "4", "6,26"])), "-E", str(rng.choice(["1", "2,1"]))]
    if rng.random() < .15: args += ["-T", "15"]
    if rng.random() < .15 and "-a" not in args and "map-pb" != preset: args += ["--qstrand"]
    if rng.random() < .2: args += ["-N", str(int(rng.choice([0, 2, 10]))), "-p", str(rng.choice(["0.5", "0.9"]))]
    if rng.random() < .2: args += ["-s", str(int(rng.choice([20, 40]))), "-m", str(int(rng.choice([20, 30]))), "-n", str(int(rng.choice([2, 3])))]
    if rng.random() < .15: args += ["-K", "1k"]
    if rng.random() < .1: args += ["-g", str(int(rng.choice([200, 1000])))]
    if nc == 2 and rng.random() < .3:  # the second contig as an ALT haplotype of part of the first
        m_ = min(len(gs[1]), 3000, len(gs[0]) - 500)
        gs[1][:m_] = gs[0][500:500 + m_]
        gs[1][rng.integers(0, m_, max(1, m_ // 100))] = list(rng.choice(list(b"ACGT"), max(1, m_ // 100)))
        synth.write_fasta(rf, ["c%d" % i for i in range(nc)], [g.tobytes() for g in gs])
        open(d + "/alt.txt", "w").write("c1\n")
        args += ["--alt", d + "/alt.txt"] + (["--alt-drop", "0.3"] if rng.random() < .5 else [])
    qarg = [rf, rf] if preset == "ava-ont" and rng.random() < .5 else [rf, qf]
    full = args + qarg
    if os.environ.get("DRY"):
        return seed, True, " ".join(full), 0, ""
    x = subprocess.run([O.REF_BIN, "-t", "2"] + full, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    y = subprocess.run([EMU, "-t", "3"] + full, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=240)
    xs = [l for l in x.stdout.decode().splitlines() if not l.startswith("@PG")]; ys = [l for l in y.stdout.decode().splitlines() if not l.startswith("@PG")]
    ok = xs == ys and (x.returncode == 0) == (y.returncode == 0)
    info = ""
    if not ok:
        refused = b"not implemented" in y.stderr or b"refusing" in y.stderr
        if x.returncode != 0 and y.returncode != 0: ok = True; info = "both fail"
        elif refused: info = "REFUSED " + y.stderr.decode()[-200:]
        else:
            diff = [(p, q) for p, q in zip(xs, ys) if p != q][:1]
            info = "rc %d/%d lines %d/%d %s ERR:%s" % (x.returncode, y.returncode, len(xs), len(ys), [(a[:250], b[:250]) for a, b in diff], y.stderr.decode()[-300:])
    return seed, ok, " ".join(args), len(xs), info

def splice_case(seed):
    """-x splice with random transcripts, canonical / non-canonical introns, optional --junc-bed / --spsc files and options"""
    rng = np.random.default_rng(seed)
    d = tempfile.mkdtemp(prefix="fzs%d_" % seed)
    g = np.frombuffer(bytes(synth.random_genome(int(rng.integers(12000, 24000)), seed)[0]), dtype=np.uint8).copy()
    reads, introns = [], []
    for i in range(int(rng.integers(1, 4))):
        pos = int(rng.integers(500, len(g) - 6000)); exons = []; rev = rng.random() < 0.5
        for k in range(int(rng.integers(2, 5))):
            el = int(rng.integers(60, 260)); exons.append((pos, pos + el)); il = int(rng.integers(80, 900))
            if rng.random() < 0.6:
                d_, a_ = (b"GT", b"AG") if not rev else (b"CT", b"AC")
                g[pos + el:pos + el + 2] = list(d_); g[pos + el + il - 2:pos + el + il] = list(a_)
            introns.append((pos + el, pos + el + il, "-" if rev else "+"))
            pos += el + il
        introns.pop()
        tr = np.concatenate([g[s_:e_] for s_, e_ in exons])
        reads.append(synth.mutate_ascii(comp[tr[::-1]] if rev else tr, rng, float(rng.choice([0.0, 0.02, 0.05]))))
    rf, qf = d + "/ref.fa", d + "/q.fa"
    synth.write_fasta(rf, ["c0"], [g.tobytes()]); synth.write_fasta(qf, ["t%d" % i for i in range(len(reads))], reads)
    args = ["-x", str(rng.choice(["splice", "splice", "splice:hq"]))] + str(rng.choice(["-c", "-c --cs", "-a", "-c --MD"])).split()
    if rng.random() < .4: args += ["-u" + str(rng.choice(["f", "r", "b", "n"]))]
    if rng.random() < .3: args += ["-C", str(int(rng.choice([0, 5, 9])))]
    if rng.random() < .2: args += ["--splice-flank=no"]
    if rng.random() < .2: args += ["-G", str(int(rng.choice([500, 5000])))]
    k = rng.random()
    if k < .35:
        bed = d + "/j.bed"
        with open(bed, "w") as f:
            for st, en, sd in introns:
                if rng.random() < .8: f.write("c0\t%d\t%d\tj\t0\t%s\n" % (st + int(rng.choice([0, 0, 0, 2])), en + int(rng.choice([0, 0, 0, -1])), sd))
        args += ["--junc-bed", bed] + (["--junc-bonus", str(int(rng.choice([5, 15])))] if rng.random() < .4 else [])
    elif k < .65:
        fn = d + "/s.txt"
        with open(fn, "w") as f:
            for st, en, sd in introns:
                dp, ap = (st, en - 1) if sd == "+" else (en - 1, st)
                f.write("c0\t%d\t%s\tD\t%d\nc0\t%d\t%s\tA\t%d\n" % (dp, sd, int(rng.integers(-3, 15)), ap, sd, int(rng.integers(-3, 15))))
            for _ in range(int(rng.integers(0, 300))):
                f.write("c0\t%d\t%s\t%s\t%d\n" % (int(rng.integers(1, len(g) - 1)), "+-"[int(rng.integers(0, 2))], "DA"[int(rng.integers(0, 2))], int(rng.integers(-12, 10))))
        args += ["--spsc", fn] + (["--spsc-scale", str(rng.choice(["0.5", "1"]))] if rng.random() < .4 else []) + (["--spsc0", str(int(rng.choice([3, 8])))] if rng.random() < .3 else [])
    full = args + [rf, qf]
    try:
        x = subprocess.run([O.REF_BIN, "-t", "2"] + full, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        y = subprocess.run([EMU, "-t", "3"] + full, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
    except subprocess.TimeoutExpired:
        return seed, True, "TIMEOUT(emulation too slow) " + " ".join(args), 0, ""
    xs = [l for l in x.stdout.decode().splitlines() if not l.startswith("@PG")]; ys = [l for l in y.stdout.decode().splitlines() if not l.startswith("@PG")]
    ok = xs == ys and x.returncode == y.returncode
    info = "" if ok else "rc %d/%d lines %d/%d %s ERR:%s" % (x.returncode, y.returncode, len(xs), len(ys), [(a[:300], b[:300]) for a, b in zip(xs, ys) if a != b][:1], y.stderr.decode()[-300:])
    return seed, ok, " ".join(a if not a.startswith(d) else os.path.basename(a) for a in args), len(xs), info


if __name__ == "__main__":
    if sys.argv[1] == "--splice":
        one = splice_case
        sys.argv.pop(1)
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    with ThreadPoolExecutor(4) as ex:
        for seed, ok, a, n, info in ex.map(one, range(lo, hi)):
            print(seed, "OK " if ok else "BAD", n, a, info, flush=True)
