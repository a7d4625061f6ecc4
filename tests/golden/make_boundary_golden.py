"""Generates tests/golden/expected/example_mt.txt and mappy_mt.json with the UNMODIFIED reference: its example.c and its Cython binding
mappy are linked against the reference library built from /root/reference (tests/boundary/build_boundary.py --ref -> oracle/_ref), then run on
the MT-human / MT-orang pair. tests/test_gpu_boundary.py runs the same callers linked against libminimap2_b200.so and compares.
Run here (needs /root/reference); the outputs are committed."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "boundary"))
DATA = os.path.join(HERE, "data")

MAPPY_SNIPPET = r'''
import json, sys, threading
sys.path.insert(0, sys.argv[1])
import mappy as mp
data = sys.argv[2]
out = {}
a = mp.Aligner(data + "/MT-human.fa", preset="map-ont", n_threads=2)
assert a
name, seq, _ = next(mp.fastx_read(data + "/MT-orang.fa"))
def hits(al, s, **kw):
    return [[h.ctg, h.ctg_len, h.r_st, h.r_en, h.q_st, h.q_en, h.strand, h.mapq, h.mlen, h.blen, h.NM, h.is_primary, h.trans_strand, h.cigar_str, kw and h.cs or "", kw and h.MD or ""] for h in al.map(s, **kw)]
out["whole"] = hits(a, seq, cs=True, MD=True)
out["pieces"] = [hits(a, seq[i:i + 3000]) for i in range(0, 15000, 2500)]
out["revcomp"] = hits(a, mp.revcomp(seq[2000:9000]))
out["seq_names"] = list(a.seq_names)
out["fetch"] = a.seq("MT_human", 100, 160)
# one Aligner shared by several threads, each with its own ThreadBuffer (minimap.h:341-348: mm_map is re-entrant)
res = [None] * 6
def work(k):
    b = mp.ThreadBuffer()
    res[k] = [[h.r_st, h.r_en, h.q_st, h.q_en, h.strand, h.mapq, h.cigar_str] for h in a.map(seq[k * 2000:k * 2000 + 4000], buf=b)]
th = [threading.Thread(target=work, args=(k,)) for k in range(6)]
[t.start() for t in th]; [t.join() for t in th]
out["threads"] = res
b = mp.Aligner(seq=seq[:8000], preset="map-ont")  # index built from a string (mm_idx_str)
out["idx_str"] = hits(b, seq[1000:5000])
print(json.dumps(out))
'''


def run_mappy(build_dir):
    p = subprocess.run([sys.executable, "-c", MAPPY_SNIPPET, build_dir, DATA], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return json.loads(p.stdout.decode().strip().splitlines()[-1])


def run_example(build_dir):
    p = subprocess.run([os.path.join(build_dir, "example"), "MT-human.fa", "MT-orang.fa"], cwd=DATA, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout.decode()


if __name__ == "__main__":
    import build_boundary
    d = build_boundary.build(ref=True)
    open(os.path.join(HERE, "expected", "example_mt.txt"), "w").write(run_example(d))
    json.dump(run_mappy(d), open(os.path.join(HERE, "expected", "mappy_mt.json"), "w"), indent=0)
    print("written")
