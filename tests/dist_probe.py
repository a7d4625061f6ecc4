"""2+ GPU probe (not a pytest; run under torchrun on a multi-GPU box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_probe.py
Rank 0 builds a small synthetic index on its GPU, minimap2_b200.dist.broadcast_index() replicates the device arrays over
NCCL, every rank maps its contiguous shard of the same reads, rank 0 gathers in input order and compares with mapping all
reads itself. Prints DIST_PROBE_OK on success."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from minimap2_b200 import api, dist as mdist
L = api._setup()
n_reads, rl = 4000, 5000
idx = None
buf = np.zeros(n_reads * rl, dtype=np.uint8)
if rank == 0:
    idx = L.mmb_synth_index(int(60e6), 4, 7, 10, 15, 14)
    L.mmb_synth_reads(idx, n_reads, rl, 5, 0.10, 0.40, 0.25, buf.ctypes.data)
t = torch.from_numpy(buf).cuda()
dist.broadcast(t, src=0)
buf = t.cpu().numpy()
idx, keep = mdist.broadcast_index(idx, rank, src=0)
al = api.Aligner(preset="map-ont", _idx=idx, n_threads=16)
al.map_opt.flag |= api.MM_F_CIGAR | api.MM_F_OUT_CG
qlens = np.full(n_reads, rl, dtype=np.int32)
names = ["r%d" % i for i in range(n_reads)]


def summarize(lo, hi):
    n_regs, regs, rep = al.map_batch_raw(buf[lo * rl:hi * rl], qlens[lo:hi], names[lo:hi])
    out = []
    for i in range(hi - lo):
        rec = []
        if regs[i]:
            arr = C.cast(C.c_void_p(int(regs[i])), C.POINTER(api.Reg1))
            for j in range(n_regs[i]):
                r = arr[j]
                rec.append((r.rid, r.rs, r.re, r.qs, r.qe, r.score, r.mapq, r.mlen, r.blen))
        out.append(rec)
    al.free_batch(n_regs, regs)
    return out


cut = mdist.shard_bounds(qlens, world)
mine = summarize(cut[rank], cut[rank + 1])
merged = mdist.gather_in_order(mine, cut, rank, world, dist)
if rank == 0:
    full = summarize(0, n_reads)
    assert merged == full, "sharded result differs from the single-GPU result"
    n_hit = sum(1 for r in full if r)
    print("DIST_PROBE_OK world=%d reads=%d mapped=%d" % (world, n_reads, n_hit), flush=True)
dist.barrier()
dist.destroy_process_group()
