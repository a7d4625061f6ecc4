#!/usr/bin/env python
"""bench.py -- aligned bases/sec of the seed-chain-extend hot path (map-ont, 10 kb reads vs a 3 Gbp reference) on B200.

  python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W  # the UNMODIFIED reference (oracle/_ref) on the host CPUs

A "step" is one pass of the whole hot path (mm_sketch -> seeds -> chain -> ksw2 -> hits, i.e. mm_map semantics with -c)
over one batch of synthetic reads. Workload = BASELINE.json configs[1]: 100k x 10 kb ONT-profile reads vs a 3 Gbp
uniform-random reference in 24 contigs, `-x map-ont -c`. Data are synthetic (device-side counter-based generator,
minimap2_b200/csrc/synth.cu); the reference arm maps the same genome (written to FASTA) and a bounded sample of the same reads.

JSON line (rank 0): value = whole-job aligned bases/s with the read bases already resident in HBM when the timed region
starts; e2e = the same metric through the minimap.h-level C-ABI call mm_map_batch() with HOST buffers (H2D of the reads
and D2H of all results inside the timed region); roofline = the dominant kernel (K3 ksw2 extd2) from CUDA events on the
launch stream; cpu_baseline = the reference's own CPU code (oracle/_ref/libminimap2_ref.so) on this box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # workload knobs (defaults = BASELINE.json configs[1]); smaller values are for development only and are reported in config
    ap.add_argument("--genome-mbp", type=float, default=3000.0)
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--read-len", type=int, default=10000)
    ap.add_argument("--threads", type=int, default=0, help="host threads for orchestration / the reference arm (0 = all cores)")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="reads in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)"""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [l.strip().split(", ") for l in open(self.path) if l.strip()]
            sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = max(float(r[1]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, nm in enumerate(names):
                if any(len(r) > 3 + k and r[3 + k].strip().lower() == "active" for r in rows):
                    out["reasons"].append(nm)
            os.unlink(self.path)
        except Exception:
            pass
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def aligned_bases(n_regs, regs, api):
    tot = 0
    for i in np.nonzero(n_regs)[0]:
        arr = C.cast(C.c_void_p(int(regs[i])), C.POINTER(api.Reg1))
        for j in range(n_regs[i]):
            if arr[j].id == arr[j].parent:
                tot += arr[j].qe - arr[j].qs
    return tot


# ---------------------------------------------------------------------------------------------------------------------
def reference_lib():
    from oracle_lib import REF_SO
    if not os.path.exists(REF_SO):
        return None
    L = C.CDLL(REF_SO)
    return L


def run_reference_sample(ref_fa, reads_fa, n_threads, steps=1, warmup=0, log=None):
    """Times the UNMODIFIED reference (libminimap2_ref.so built from /root/reference by oracle/Makefile) through its own
    public API: mm_idx_reader_read (index build, untimed) then mm_map_file per step on the sample file. Returns
    (aligned bases per step, [seconds per timed step], index seconds)."""
    from minimap2_b200 import api  # only the ctypes struct mirrors (IdxOpt/MapOpt share the reference's layout)
    L = reference_lib()
    if L is None:
        return None
    L.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(api.IdxOpt), C.POINTER(api.MapOpt)]
    L.mm_idx_reader_open.restype = C.c_void_p
    L.mm_idx_reader_open.argtypes = [C.c_char_p, C.POINTER(api.IdxOpt), C.c_char_p]
    L.mm_idx_reader_read.restype = C.c_void_p
    L.mm_idx_reader_read.argtypes = [C.c_void_p, C.c_int]
    L.mm_idx_reader_close.argtypes = [C.c_void_p]
    L.mm_mapopt_update.argtypes = [C.POINTER(api.MapOpt), C.c_void_p]
    L.mm_map_file.restype = C.c_int
    L.mm_map_file.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(api.MapOpt), C.c_int]
    L.mm_idx_destroy.argtypes = [C.c_void_p]
    C.c_int.in_dll(L, "mm_verbose").value = 1
    io, mo = api.IdxOpt(), api.MapOpt()
    L.mm_set_opt(None, C.byref(io), C.byref(mo))
    L.mm_set_opt(b"map-ont", C.byref(io), C.byref(mo))
    mo.flag |= 0x004 | 0x020  # -c
    t0 = time.time()
    rdr = L.mm_idx_reader_open(ref_fa.encode(), C.byref(io), None)
    mi = L.mm_idx_reader_read(rdr, n_threads)
    L.mm_idx_reader_close(rdr)
    L.mm_mapopt_update(C.byref(mo), mi)
    t_idx = time.time() - t0
    if log:
        log("reference index built in %.1fs (mid_occ=%d)" % (t_idx, mo.mid_occ))
    times, bases = [], 0
    out_path = reads_fa + ".ref.paf"
    for it in range(warmup + steps):
        sys.stdout.flush()
        saved = os.dup(1)
        fd = os.open(out_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        os.dup2(fd, 1)
        t = time.time()
        L.mm_map_file(mi, reads_fa.encode(), C.byref(mo), n_threads)
        libc = C.CDLL(None)
        libc.fflush(None)
        dt = time.time() - t
        os.dup2(saved, 1)
        os.close(fd); os.close(saved)
        if it >= warmup:
            times.append(dt)
    bases = 0
    with open(out_path) as f:
        for line in f:
            c = line.split("\t", 13)
            if "tp:A:P" in line:
                bases += int(c[3]) - int(c[2])
    L.mm_idx_destroy(mi)
    return bases, times, t_idx


# ---------------------------------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_threads_all = os.cpu_count() or 1
    nthr = a.threads or max(1, n_threads_all // max(1, world))
    cfg = {"workload": "map-ont: %d synthetic %d bp ONT-profile reads (10%% err, sub/ins/del 40/25/35) vs synthetic %.0f Mbp reference "
                       "(24 contigs), -x map-ont -c" % (a.reads, a.read_len, a.genome_mbp),
           "preset": "map-ont", "k": 15, "w": 10, "reads_per_step_per_gpu": a.reads, "read_len": a.read_len,
           "genome_mbp": a.genome_mbp, "host_threads": nthr,
           "l2_policy": "inputs larger than L2 (>=1 GB of read bases + multi-GB index touched every step)"}

    def log(msg):
        if rank == 0:
            print("[bench] " + msg, file=sys.stderr, flush=True)

    os.environ["MM_B200_DEVICE"] = str(local_rank)
    import torch
    dist_on = world > 1
    if a.impl == "reference":
        if rank != 0:
            return 0
        import minimap2_b200 as mb
        from minimap2_b200 import api
        L = api._setup()
        tmp = tempfile.mkdtemp(prefix="mm2bench_")
        log("generating the synthetic genome/reads for the reference arm")
        idx = L.mmb_synth_index(int(a.genome_mbp * 1e6), 24, 11, 10, 15, 14)
        ref_fa, reads_fa = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "sample.fa")
        L.mmb_idx_write_fasta(idx, ref_fa.encode())
        ns = min(a.cpu_sample, a.reads)
        buf = np.zeros(ns * a.read_len, dtype=np.uint8)
        L.mmb_synth_reads(idx, ns, a.read_len, 12, 0.10, 0.40, 0.25, buf.ctypes.data)
        with open(reads_fa, "wb") as f:
            for i in range(ns):
                f.write(b">r%d\n" % i); f.write(buf[i * a.read_len:(i + 1) * a.read_len].tobytes()); f.write(b"\n")
        L.mm_idx_destroy(idx)
        res = run_reference_sample(ref_fa, reads_fa, n_threads_all, steps=a.steps, warmup=a.warmup, log=log)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libminimap2_ref.so missing"}))
            return 0
        bases, times, t_idx = res
        tot_t = sum(times)
        val = bases * len(times) / tot_t
        line = {"metric": "aligned bases/sec (map-ont, 10 kb reads)", "value": val, "unit": "bases/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8/int32", "data": "synthetic", "impl": "reference", "config": dict(cfg, sample_reads=ns, cpu_threads=n_threads_all),
                "cpu_baseline": {"value": val, "unit": "bases/s", "cores": n_threads_all, "kind": "reference",
                                 "sample": "%d of the %d reads per step, same 3 Gbp reference; mm_map_file() wall time, index build (%.0fs) excluded" % (ns, a.reads, t_idx)},
                "e2e": {"value": val, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------- this repo's arm ----------------
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    # host worker pool per rank: half of this rank's share of the logical CPUs (the group threads that feed the GPU need idle cores)
    os.environ.setdefault("MM_B200_HOST_THREADS", str(max(8, n_threads_all // (2 * max(1, world)))))
    import minimap2_b200 as mb
    from minimap2_b200 import api
    L = api._setup()
    t0 = time.time()
    # Every rank builds the same index deterministically on its own GPU from the seed (the device build takes seconds).
    # A loaded (non-synthetic) index is broadcast instead: see minimap2_b200/dist.py (NCCL broadcast of the device arrays).
    idx = L.mmb_synth_index(int(a.genome_mbp * 1e6), 24, 11, 10, 15, 14)
    log("index built on device in %.1fs" % (time.time() - t0))
    al = api.Aligner(preset="map-ont", _idx=idx, n_threads=nthr)
    al.map_opt.flag |= api.MM_F_CIGAR | api.MM_F_OUT_CG
    buf = np.zeros(a.reads * a.read_len, dtype=np.uint8)
    L.mmb_synth_reads(idx, a.reads, a.read_len, 12 + 1000 * rank, 0.10, 0.40, 0.25, buf.ctypes.data)
    qlens = np.full(a.reads, a.read_len, dtype=np.int32)
    names = ["r%d" % i for i in range(a.reads)]
    ctx = L.mmb_default_ctx_c()
    prepared = al.prepare_batch(buf, qlens, names)

    def barrier():
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, resident):
        L.mmb_set_resident_reads(1 if resident else 0)
        bases, times = 0, []
        for _ in range(n):
            t = time.perf_counter()
            n_regs, regs, rep = al.map_prepared(prepared)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t)
            bases = aligned_bases(n_regs, regs, api)
            al.free_batch(n_regs, regs)
        return bases, times

    log("warm-up x%d" % a.warmup)
    run_steps(max(a.warmup, 3), True)
    L.mmb_launch_count_all(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # --- timed region A: `value` (read bases resident in HBM; only the mm_map_batch calls are timed) ---
    barrier()
    bases, times_a = run_steps(a.steps, True)
    barrier()
    t_a = sum(times_a)
    launches = int(L.mmb_launch_count_all(0))
    # --- timed region B: `e2e` (host buffers in, results out) ---
    barrier()
    bases_b, times_b = run_steps(a.steps, False)
    barrier()
    t_b = sum(times_b)
    clocks = sampler.stop()
    # --- per-kernel device time for the roofline: one extra step with the read groups serialised (one stream), so that
    #     CUDA-event durations are not inflated by kernels of other groups sharing the SMs ---
    L.mmb_set_groups(-int(os.environ.get("MM_B200_GROUPS", "12")))  # the default group count, run one after another
    L.mmb_profile_enable_all(1)
    for k in range(6):
        L.mmb_profile_ms_all(k, 1); L.mmb_profile_units_all(k, 1); L.mmb_profile_bytes_all(k, 1); L.mmb_profile_scopes_all(k, 1)
    n_prof = 2
    run_steps(n_prof, True)
    prof = {}
    for k, nm in enumerate(["sketch", "seed", "sort", "chain", "ksw", "other"]):
        prof[nm] = {"ms": L.mmb_profile_ms_all(k, 0), "units": int(L.mmb_profile_units_all(k, 0)), "bytes": int(L.mmb_profile_bytes_all(k, 0)),
                    "scopes": int(L.mmb_profile_scopes_all(k, 0))}
    L.mmb_profile_enable_all(0)
    L.mmb_set_groups(0)
    # max over ranks
    tt = torch.tensor([t_a, t_b], dtype=torch.float64, device="cuda")
    bb = torch.tensor([float(bases), float(bases_b)], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(bb, op=dist.ReduceOp.SUM)
    t_a, t_b = float(tt[0]), float(tt[1])
    tot_bases_a, tot_bases_b = float(bb[0]), float(bb[1])
    value = tot_bases_a * a.steps / t_a
    e2e = tot_bases_b * a.steps / t_b
    if rank != 0:
        if dist_on:
            dist.destroy_process_group()
        return 0
    # --- roofline of the dominant kernel (K3) ---
    peak, peak_src = measured_peaks()
    k = prof["ksw"]
    n_launch_ksw = max(1, k["scopes"])
    ksw_gbs = (k["bytes"] / 1e9) / (k["ms"] / 1e3) if k["ms"] > 0 else 0.0
    roofline = {"kernel": "K3 ksw2 kernels (ksw_pk_kernel + ksw_extd2_kernel)", "bound": "hbm", "achieved": ksw_gbs, "peak": peak, "unit": "GB/s", "frac": ksw_gbs / peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": k["bytes"] / n_launch_ksw, "avg_launch_ms": k["ms"] / n_launch_ksw,
                "gcups": (k["units"] / 1e9) / (k["ms"] / 1e3) if k["ms"] > 0 else 0.0,
                "note": "ALU-bound by construction: ~%.0f DP cells per read base at 1 B/cell traceback; HBM fraction is expected to be small (SURVEY 8d)" % (k["units"] / max(1.0, tot_bases_a / max(1, world) * n_prof)),
                "stage_ms_per_step": {nm: prof[nm]["ms"] / n_prof for nm in prof},
                "timing": "CUDA events on the launch stream, %d profiled steps with the scheduler's read groups serialised" % n_prof}
    try:  # per-stage algorithmic bandwidth (DESIGN.md section 3 definitions); informational, never allowed to break the line
        roofline["stages"] = {nm: {"ms_per_step": prof[nm]["ms"] / n_prof, "algorithmic_gb_per_step": prof[nm]["bytes"] / n_prof / 1e9,
                                   "gb_per_s": (prof[nm]["bytes"] / 1e9) / (prof[nm]["ms"] / 1e3) if prof[nm]["ms"] > 0 else 0.0,
                                   "frac_of_hbm_peak": ((prof[nm]["bytes"] / 1e9) / (prof[nm]["ms"] / 1e3) / peak) if prof[nm]["ms"] > 0 and peak else 0.0}
                              for nm in ("sketch", "seed", "sort", "chain", "ksw")}
    except Exception:
        pass
    tp = os.path.join(ROOT, "profiles", "ksw_traffic.json")
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {"metric": "aligned bases/sec (map-ont, 10 kb reads)", "value": value, "unit": "bases/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_a / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 DP cells (int32 scores), u64 hashes, f32 chain penalties", "data": "synthetic", "config": cfg,
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "bases/s", "ms_per_step": 1e3 * t_b / a.steps,
                    "h2d_bytes_per_step": int(a.reads * a.read_len + 12 * a.reads), "d2h_bytes_per_step": int(L.mmb_last_d2h_bytes())},
            "roofline": roofline}
    # --- CPU baseline: the reference's own code on this box's cores, bounded sample ---
    if not a.no_cpu_baseline:
        try:
            tmp = tempfile.mkdtemp(prefix="mm2bench_")
            ref_fa, reads_fa = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "sample.fa")
            log("writing the reference FASTA for the CPU baseline")
            L.mmb_idx_write_fasta(idx, ref_fa.encode())
            ns = min(a.cpu_sample, a.reads)
            with open(reads_fa, "wb") as f:
                for i in range(ns):
                    f.write(b">r%d\n" % i); f.write(buf[i * a.read_len:(i + 1) * a.read_len].tobytes()); f.write(b"\n")
            al.close()
            idx = None
            res = run_reference_sample(ref_fa, reads_fa, n_threads_all, steps=1, warmup=0, log=log)
            if res is not None:
                cb, ct, t_idx = res
                line["cpu_baseline"] = {"value": cb / ct[0], "unit": "bases/s", "cores": n_threads_all, "kind": "reference",
                                        "sample": "%d of the %d reads of one step vs the same reference; mm_map_file() wall %.2fs (index build %.0fs excluded)" % (ns, a.reads, ct[0], t_idx)}
            for fn in (ref_fa, reads_fa, reads_fa + ".ref.paf"):
                if os.path.exists(fn):
                    os.unlink(fn)
        except Exception as e:  # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": n_threads_all, "kind": "reference", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
