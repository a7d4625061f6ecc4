"""Ad-hoc scale probe (not a pytest): python tests/scale_probe.py <genome_mbp> <n_reads> [read_len]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import minimap2_b200 as mb
from minimap2_b200 import api

mbp = float(sys.argv[1]); n_reads = int(sys.argv[2]); rl = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
nthr = int(os.environ.get("MM_THREADS", "64"))
L = api._setup()
t0 = time.time()
idx = L.mmb_synth_index(int(mbp * 1e6), 24, 11, 10, 15, 14)
t1 = time.time()
print("index: %.2fs for %.0f Mbp" % (t1 - t0, mbp), flush=True)
al = api.Aligner(preset="map-ont", _idx=idx, n_threads=nthr)
if not os.environ.get('NO_CIGAR'):
    al.map_opt.flag |= api.MM_F_CIGAR | api.MM_F_OUT_CG
print("mid_occ", al.map_opt.mid_occ, flush=True)
buf = np.zeros(n_reads * rl, dtype=np.uint8)
L.mmb_synth_reads(idx, n_reads, rl, 12, 0.10, 0.40, 0.25, buf.ctypes.data)
t2 = time.time()
print("reads: %.2fs" % (t2 - t1), flush=True)
qlens = np.full(n_reads, rl, dtype=np.int32)
ctx = L.mmb_default_ctx_c()
names = ["r%d" % i for i in range(n_reads)]
sweep = os.environ.get("SWEEP", "")
cfgs = [tuple(int(x) for x in c.split(":")) for c in sweep.split(",")] if sweep else [None] * 3
for it, cfg in enumerate(cfgs):
    if cfg is not None:
        L.mmb_set_groups(cfg[0]); L.mmb_set_gpu_slots(cfg[1])
        print("groups %d slots %d" % cfg, flush=True)
    L.mmb_profile_enable_all(1 if (it == len(cfgs) - 1 and (not sweep or cfg[0] < 0)) else 0)
    t = time.time(); c0 = time.process_time()
    n_regs, regs, rep = al.map_batch_raw(buf, qlens, names)
    dt = time.time() - t; print('  cpu seconds %.2f' % (time.process_time() - c0))
    # aligned bases over primary records
    tot = 0
    for i in range(0, n_reads):
        if regs[i]:
            arr = C.cast(C.c_void_p(int(regs[i])), C.POINTER(api.Reg1))
            for j in range(n_regs[i]):
                if arr[j].id == arr[j].parent:
                    tot += arr[j].qe - arr[j].qs
    print("iter %d: %.3fs  mapped reads %d/%d aligned bases %d  -> %.1f Mbases/s" % (it, dt, int((n_regs > 0).sum()), n_reads, tot, tot / dt / 1e6), flush=True)
    al.free_batch(n_regs, regs)
names_k = ["sketch", "seed", "sort", "chain", "ksw", "other"]
for i, nm in enumerate(names_k):
    print("  %-7s %.2f ms  units %d" % (nm, L.mmb_profile_ms_all(i, 0), L.mmb_profile_units_all(i, 0)))
print("launches", L.mmb_launch_count_all(0))
