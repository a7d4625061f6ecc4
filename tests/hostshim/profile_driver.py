"""Host-side cost of the alignment driver (csrc/align.cc) per read, on the CPU: the replayable mm_align_skeleton restatement is run
through tests/hostshim/alignshim.cc (oracle ksw2 as the job executor) on synthetic map-ont reads, and the cycles of the requesting
replay, the final replay and the profiled sections are printed. Not a test: a development tool for the host phases that the GPU
scheduler overlaps with kernels (stage 3). Usage: python tests/hostshim/profile_driver.py [n_reads] [read_len]"""
import ctypes as C
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle_lib as O
import synth
import test_aligndriver_vs_ref as T


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rlen = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    H, R, api = T.libs.__wrapped__() if hasattr(T.libs, "__wrapped__") else T.libs.__pytest_wrapped__.obj()
    contigs = synth.random_genome(2_000_000, 17, n_contigs=2, repeat_frac=0.05)
    reads = synth.make_reads(contigs, n_reads, rlen, 0.10, 117, chimeric_frac=0.0)
    mi, keep = T.build_ref_index(R, contigs, ["chr0", "chr1"])
    io, mo = api.IdxOpt(), api.MapOpt()
    R.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(api.IdxOpt), C.POINTER(api.MapOpt)]
    R.mm_set_opt(None, C.byref(io), C.byref(mo)); R.mm_set_opt(b"map-ont", C.byref(io), C.byref(mo))
    mo.flag |= api.MM_F_CIGAR
    mo.mid_occ = 50
    R.mm_mapopt_update(C.byref(mo), mi)
    oidx = O.OracleIndex([bytes(c) for c in contigs], ["chr0", "chr1"], 10, 15)
    pg = float(np.float32(mo.chain_gap_scale * 0.01 * 15)); ps = float(np.float32(mo.chain_skip_scale * 0.01 * 15))
    work = []
    for qi, rd in enumerate(reads):
        qstr = bytes(rd)
        a, rep, mini = oidx.anchors(qstr, mid_occ=mo.mid_occ, q_occ_frac=mo.q_occ_frac, max_max_occ=mo.max_max_occ, occ_dist=mo.occ_dist)
        if len(a) == 0:
            continue
        u, b = O.ref_lchain_dp(a, mo.max_gap, mo.max_gap, mo.bw, mo.max_chain_skip, mo.max_chain_iter, mo.min_cnt, mo.min_chain_score, pg, ps, 0)
        if len(u) == 0:
            continue
        n = len(u); uu = u.copy(); bb = np.ascontiguousarray(b.copy())
        regs0 = R.mm_gen_regs(None, C.c_uint32(12345 + qi), len(qstr), n, uu.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p), 0)
        R.mm_set_parent(None, C.c_float(mo.mask_level), mo.mask_len, n, C.c_void_p(regs0), mo.a * 2 + mo.b, 0, C.c_float(mo.alt_drop))
        nn = C.c_int(n)
        R.mm_select_sub(None, C.c_float(mo.pri_ratio), 30, mo.best_n, 1, int(mo.max_gap * 0.8), C.byref(nn), C.c_void_p(regs0))
        n0 = R.mm_filter_strand_retained(nn.value, C.c_void_p(regs0))
        work.append((qstr, C.string_at(regs0, n0 * T.REG_SIZE), n0, bb))
    oidx.close()
    ns = H.hs_prof_n_sections()
    H.hs_prof_enable(1)
    t0 = time.time()
    for qstr, snap, n0, bb in work:
        nm = C.c_int(n0); waves = C.c_int(0)
        a_mine = np.ascontiguousarray(bb.copy())
        pm = H.hs_align_skeleton(C.byref(mo), mi, len(qstr), qstr, C.byref(nm), snap, len(a_mine), a_mine.ctypes.data_as(C.c_void_p), C.byref(waves))
        H.hs_free_regs(nm.value, C.c_void_p(pm))
    wall = time.time() - t0
    sec = (C.c_uint64 * ns)(); rep = (C.c_uint64 * 2)()
    H.hs_prof_read(sec, rep)
    # TSC rate

    names = ["SKEL", "TSEQ", "ZDROP", "EXTRA", "FETCH", "APPEND", "PRE", "POST", "HITS"]
    ghz = float(os.environ.get("TSC_GHZ", "2.0"))
    nr = len(work)
    print("reads %d x %d bp; wall incl. oracle ksw %.1f s" % (nr, rlen, wall))
    print("requesting replay: %.1f us/read   final replay: %.1f us/read (TSC %.1f GHz assumed)" % (rep[0] / ghz / 1e3 / nr, rep[1] / ghz / 1e3 / nr, ghz))
    for i in range(ns):
        print("  %-7s %.1f us/read (all replays)" % (names[i] if i < len(names) else str(i), sec[i] / ghz / 1e3 / nr))


if __name__ == "__main__":
    main()
