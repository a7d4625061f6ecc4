"""K3 micro-benchmark (not a pytest): map-ont-shaped ksw2 jobs through the C-ABI with host buffers; prints GCUPS from CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ctypes as C
import numpy as np
import minimap2_b200._lib as _L
if os.environ.get('AB_LIB'):
    _L.LIB_PATH = os.environ['AB_LIB']  # A/B of two builds on the same box (development only)
import minimap2_b200 as mb
from minimap2_b200 import kernels as K
from minimap2_b200._lib import lib, KswJob, KswRes
import oracle_lib as O
import synth

n_jobs = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
flag = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x08
rng = np.random.default_rng(1)
ctx = mb.Context(0)
L = lib()
G = 240 * n_jobs + 1000
t_all = synth.ALPHA[rng.integers(0, 4, G)]
q_all = synth.mutate_ascii(t_all, rng, 0.10)
lut = np.full(256, 4, dtype=np.uint8); lut[65] = 0; lut[67] = 1; lut[71] = 2; lut[84] = 3
t4 = lut[t_all]; q4 = lut[q_all]
jobs = (KswJob * n_jobs)()
ratio = len(q4) / len(t4)
cells = 0
tl_lo, tl_hi = (int(x) for x in os.environ.get('KSW_TLEN', '200,270').split(','))
qfrac = float(os.environ.get('KSW_QFRAC', '1.0'))  # extension-shaped jobs: KSW_TLEN=60,110 KSW_QFRAC=0.5 with flag 0x40
tls = rng.integers(tl_lo, tl_hi, n_jobs)
for i in range(n_jobs):
    tl = int(tls[i]); ts = i * 240
    qs = int(ts * ratio); ql = max(1, int(tl * ratio * qfrac))
    j = jobs[i]
    j.q_start, j.t_start, j.q_step, j.t_step, j.qlen, j.tlen = qs, ts, 1, 1, ql, tl
    j.w, j.zdrop, j.end_bonus, j.flag = (751, 400, 10, flag) if flag & 0x40 else (30001, 400, -1, flag)
    cells += ql * tl
res = (KswRes * n_jobs)()
cig = np.zeros(n_jobs * 300, dtype=np.uint32)
sc = K.make_score(O.simple_mat(2, 4, 1), 4, 2, 24, 1)
for it in range(3):
    L.mmb_profile_enable(ctx.h, 1)
    L.mmb_profile_ms(ctx.h, 4, 1); L.mmb_profile_units(ctx.h, 4, 1)
    t = time.time()
    used = L.mmb_ksw_batch_host(ctx.h, C.byref(sc), n_jobs, jobs, q4.ctypes.data, len(q4), t4.ctypes.data, len(t4), res, cig.ctypes.data, len(cig))
    dt = time.time() - t
    ms = L.mmb_profile_ms(ctx.h, 4, 0)
    print("iter %d: wall %.3fs kernel %.2f ms cells %.3g -> %.1f GCUPS (used cigar %d, score[0]=%d)" % (it, dt, ms, cells, cells / ms / 1e6, used, res[0].score), flush=True)
