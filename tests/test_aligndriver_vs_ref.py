"""CPU: the product's alignment driver (minimap2_b200/csrc/align.cc: mm_align_skeleton / mm_align1 / mm_align1_inv restated as a
replayable routine) against the reference's mm_align_skeleton on identical chains. The ksw2 jobs the driver requests are
executed by the oracle's ksw2 (tests/hostshim/alignshim.cc plays the GPU's role), so this checks the driver logic -- window
selection, gap-fill two-pass rule, z-drop splits, inversion probes, CIGAR assembly, statistics, filters -- without a GPU and
independently of the CUDA kernels. Compared: every field of every mm_reg1_t and of its mm_extra_t, CIGAR included."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import oracle_lib as O
import synth
from test_hostlogic_vs_ref import Idx, REG_SIZE

ROOT = O.ROOT
SHIM = os.path.join(ROOT, "tests", "hostshim", "_build", "libalignshim.so")
pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def libs():
    from minimap2_b200 import api
    O.build_oracle() if hasattr(O, "build_oracle") else None
    os.makedirs(os.path.dirname(SHIM), exist_ok=True)
    csrc = os.path.join(ROOT, "minimap2_b200", "csrc")
    src = [os.path.join(csrc, "align.cc"), os.path.join(csrc, "hits.cc"), os.path.join(csrc, "format.cc"), os.path.join(ROOT, "tests", "hostshim", "hostshim.cc"),
           os.path.join(ROOT, "tests", "hostshim", "alignshim.cc")]
    deps = src + [os.path.join(csrc, h) for h in ("hostlogic.h", "annot.h", "mm_algo.cuh")] + [os.path.join(ROOT, "include", "mm_b200.h")]
    if not os.path.exists(SHIM) or any(os.path.getmtime(s) > os.path.getmtime(SHIM) for s in deps):
        inc = ["-I" + os.path.join(ROOT, "include"), "-I" + csrc, "-I/usr/local/cuda/include"]
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared"] + inc + src +
                              ["-L" + O.ORACLE_DIR, "-lmm2oracle", "-Wl,-rpath," + O.ORACLE_DIR, "-o", SHIM])
    H = C.CDLL(SHIM)
    R = O.ref()
    H.hs_align_skeleton.restype = C.c_void_p
    R.mm_align_skeleton.restype = C.c_void_p
    R.mm_gen_regs.restype = C.c_void_p
    R.mm_idx_str.restype = C.POINTER(Idx)
    return H, R, api


def build_ref_index(R, contigs, names, w=10, k=15):
    n = len(contigs)
    seqs = (C.c_char_p * n)(*[bytes(c) for c in contigs])
    nms = (C.c_char_p * n)(*[x.encode() for x in names])
    return R.mm_idx_str(w, k, 0, 14, n, seqs, nms), (seqs, nms)


def extra_of(reg_ptr, i):
    """(header ints, cigar list) of regs[i].p, or None"""
    p = C.cast(C.c_void_p(reg_ptr + i * REG_SIZE + 72), C.POINTER(C.c_void_p))[0]
    if not p:
        return None
    hdr = (C.c_uint32 * 7).from_address(p)
    n_cigar = hdr[6]
    cig = (C.c_uint32 * n_cigar).from_address(p + 28)
    return (tuple(hdr[1:7]), list(cig))


class KString(C.Structure):
    _fields_ = [("l", C.c_size_t), ("m", C.c_size_t), ("s", C.c_void_p)]


class BSeq1(C.Structure):  # mm_bseq1_t (bseq.h:13-17)
    _fields_ = [("l_seq", C.c_int), ("rid", C.c_int), ("name", C.c_char_p), ("seq", C.c_char_p), ("qual", C.c_char_p), ("comment", C.c_char_p)]


F_OUT_SAM, F_OUT_CG, F_OUT_CS, F_OUT_MD, F_OUT_CS_LONG = 0x008, 0x020, 0x040, 0x1000000, 0x800  # minimap.h:26-47


def compare_formats(H, R, api, mi, qname, qstr, n, pm, pr, rep_len, flag):
    """format.cc vs format.c on the aligned hits: PAF with cg/cs/MD tags and SAM records (incl. SA tags, clipping)"""
    t = BSeq1(len(qstr), 0, qname.encode(), qstr, None, None)
    buf = C.create_string_buffer(1 << 20)
    for extra in (0, F_OUT_CG, F_OUT_CG | F_OUT_CS | F_OUT_MD, F_OUT_CS | F_OUT_CS_LONG):
        fl = (flag | extra)
        for j in range(n):
            ks = KString(0, 0, None)
            R.mm_write_paf3(C.byref(ks), mi, C.byref(t), C.c_void_p(pr + j * REG_SIZE), None, C.c_int64(fl), C.c_int(rep_len))
            ref_line = C.string_at(ks.s, ks.l)
            R.refshim_free(C.c_void_p(ks.s))
            ln = H.hs_write_paf(buf, len(buf), mi, qname.encode(), qstr, len(qstr), C.c_void_p(pm + j * REG_SIZE), C.c_int64(fl), C.c_int(rep_len))
            assert ln >= 0 and buf.raw[:ln] == ref_line, (qname, j, hex(extra), ref_line[:300], buf.raw[:min(ln, 300)])
    if flag & 0x100000000:  # SAM is not defined in query-strand mode (mm_check_opt, options.c)
        return
    fl = flag | F_OUT_SAM | F_OUT_MD
    n_arr = (C.c_int * 1)(n); regs_arr = (C.c_void_p * 1)(pr)
    for j in range(n):
        ks = KString(0, 0, None)
        R.mm_write_sam3(C.byref(ks), mi, C.byref(t), 0, j, 1, n_arr, regs_arr, None, C.c_int64(fl), C.c_int(rep_len))
        ref_line = C.string_at(ks.s, ks.l)
        R.refshim_free(C.c_void_p(ks.s))
        ln = H.hs_write_sam(buf, len(buf), mi, qname.encode(), qstr, None, len(qstr), j, n, C.c_void_p(pm), C.c_int64(fl), C.c_int(rep_len))
        assert ln >= 0 and buf.raw[:ln] == ref_line, (qname, j, ref_line[:300], buf.raw[:min(ln, 300)])


def run_case(H, R, api, contigs, names, reads, preset="map-ont", w=10, is_cdna=0, k=15, tweak=None, bed=None, spsc=None):
    mi, keep = build_ref_index(R, contigs, names, w=w, k=k)
    if bed:  # junction annotation (main.c:467-471): both drivers read it from mi->I
        R.mm_idx_bed_read(mi, bed.encode(), 1)
    if spsc:  # splice scores (main.c:482-486): mi->spsc, max bonus as main.c computes it for the splice preset (q2=32, q=2)
        R.mm_idx_spsc_read2.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_float]
        R.mm_idx_spsc_read2(mi, spsc.encode(), 30, 0.7)
    io, mo = api.IdxOpt(), api.MapOpt()
    R.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(api.IdxOpt), C.POINTER(api.MapOpt)]
    R.mm_set_opt(None, C.byref(io), C.byref(mo)); R.mm_set_opt(preset.encode(), C.byref(io), C.byref(mo))
    mo.flag |= api.MM_F_CIGAR
    mo.mid_occ = 50
    if tweak:
        tweak(mo)
    R.mm_mapopt_update(C.byref(mo), mi)
    oidx = O.OracleIndex([bytes(c) for c in contigs], names, w, k)
    pg = float(np.float32(mo.chain_gap_scale * 0.01 * k))
    n_checked = n_split = 0
    for qi, rd in enumerate(reads):
        qstr = bytes(rd); qlen = len(qstr)
        qst = 1 if mo.flag & 0x100000000 else 0  # MM_F_QSTRAND: reverse hits in the coordinates of the target's other strand (map.c:188-192)
        a, rep, mini = oidx.anchors(qstr, flag=(0x100000000 if qst else 0), mid_occ=mo.mid_occ, q_occ_frac=mo.q_occ_frac, max_max_occ=mo.max_max_occ, occ_dist=mo.occ_dist)
        if len(a) == 0:
            continue
        gap_ref = mo.max_gap_ref if mo.max_gap_ref > 0 else mo.max_gap  # map.c:262-269
        ps = float(np.float32(mo.chain_skip_scale * 0.01 * k))
        u, b = O.ref_lchain_dp(a, gap_ref, mo.max_gap, mo.bw, mo.max_chain_skip, mo.max_chain_iter, mo.min_cnt, mo.min_chain_score, pg, ps, is_cdna)
        if len(u) == 0:
            continue
        n = len(u)
        uu = u.copy(); bb = np.ascontiguousarray(b.copy())
        hash_ = 12345 + qi
        regs0 = R.mm_gen_regs(None, C.c_uint32(hash_), qlen, n, uu.ctypes.data_as(C.c_void_p), bb.ctypes.data_as(C.c_void_p), qst)
        R.mm_set_parent(None, C.c_float(mo.mask_level), mo.mask_len, n, C.c_void_p(regs0), mo.a * 2 + mo.b, 0, C.c_float(mo.alt_drop))
        nn = C.c_int(n)
        R.mm_select_sub(None, C.c_float(mo.pri_ratio), k * 2, mo.best_n, 1, int(mo.max_gap * 0.8), C.byref(nn), C.c_void_p(regs0))
        n0 = nn.value if qst else R.mm_filter_strand_retained(nn.value, C.c_void_p(regs0))  # map.c:333-336
        snap = C.string_at(regs0, n0 * REG_SIZE)
        # mine first (inputs are const), then the reference (consumes regs0 and rewrites the anchors)
        nm = C.c_int(n0); waves = C.c_int(0)
        a_mine = np.ascontiguousarray(bb.copy())
        pm = H.hs_align_skeleton(C.byref(mo), mi, qlen, qstr, C.byref(nm), snap, len(a_mine), a_mine.ctypes.data_as(C.c_void_p), C.byref(waves))
        nr = C.c_int(n0)
        a_ref = np.ascontiguousarray(bb.copy())
        pr = R.mm_align_skeleton(None, C.byref(mo), mi, qlen, qstr, C.byref(nr), C.c_void_p(regs0), a_ref.ctypes.data_as(C.c_void_p))
        assert nm.value == nr.value, (qi, nm.value, nr.value)
        for i in range(nr.value):
            x = C.string_at(pm + i * REG_SIZE, 72); y = C.string_at(pr + i * REG_SIZE, 72)
            assert x == y, (qi, i, np.frombuffer(x, dtype=np.int32), np.frombuffer(y, dtype=np.int32))
            assert extra_of(pm, i) == extra_of(pr, i), (qi, i)
        compare_formats(H, R, api, mi, "read%d" % qi, qstr, nr.value, pm, pr, rep, mo.flag)
        n_checked += 1
        n_split += nr.value > n0
        H.hs_free_regs(nm.value, C.c_void_p(pm))
    oidx.close()
    return n_checked, n_split


def test_driver_on_synthetic_ont(libs):
    H, R, api = libs
    contigs = synth.random_genome(250_000, 17, n_contigs=2, repeat_frac=0.15)
    reads = synth.make_reads(contigs, 50, 3000, 0.10, 117, chimeric_frac=0.15)
    n, _ = run_case(H, R, api, contigs, ["chr0", "chr1"], reads)
    assert n >= 40


def test_driver_on_inversion_pair(libs):
    """test/t-inv.fa + q-inv.fa of the reference: z-drop split and mm_align1_inv (the ksw_ll_i16 probe)"""
    H, R, api = libs

    def fa(path):
        seqs, names = [], []
        for l in open(path):
            if l.startswith(">"):
                names.append(l[1:].split()[0]); seqs.append([])
            else:
                seqs[-1].append(l.strip())
        return names, [("".join(s)).encode() for s in seqs]
    data = os.path.join(ROOT, "tests", "golden", "data")
    tn, ts = fa(os.path.join(data, "t-inv.fa")); qn, qs = fa(os.path.join(data, "q-inv.fa"))
    n, _ = run_case(H, R, api, ts, tn, qs)
    assert n == len(qs)


def test_driver_on_structural_variants(libs):
    """reads with junk insertions, long deletions and inverted segments: z-drop splits (mm_test_zdrop, second exact pass), region
    splitting and the inversion probe on synthetic data"""
    H, R, api = libs
    rng = np.random.default_rng(5)
    contigs = synth.random_genome(200_000, 23, n_contigs=1, repeat_frac=0.05)
    g = np.frombuffer(bytes(contigs[0]), dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    reads = []
    for i in range(40):
        s = int(rng.integers(1000, len(g) - 12000))
        left = g[s:s + 2500]
        kind = i % 4
        if kind == 0:    # junk insertion
            mid = synth.ALPHA[rng.integers(0, 4, int(rng.integers(300, 900)))]; right = g[s + 2500:s + 5000]
        elif kind == 1:  # long deletion in the read (reference has extra sequence)
            mid = np.zeros(0, dtype=np.uint8); right = g[s + 2500 + int(rng.integers(1500, 6000)):][:2500]
        elif kind == 2:  # inverted segment
            seg = g[s + 2500:s + 2500 + int(rng.integers(300, 1200))]; mid = comp[seg[::-1]]; right = g[s + 2500 + len(seg):][:2500]
        else:            # plain
            mid = np.zeros(0, dtype=np.uint8); right = g[s + 2500:s + 5000]
        rd = np.concatenate([left, mid, right])
        reads.append(synth.mutate_ascii(rd, rng, 0.06))
    n, n_split = run_case(H, R, api, contigs, ["chr0"], reads)
    assert n >= 35 and n_split >= 3, (n, n_split)


def test_driver_on_spliced_reads(libs):
    """-x splice host logic (boundary-anchor probe, splice-model flags, both transcript strands, trans_strand, N operations in the
    statistics and in PAF/SAM) with the oracle's ksw_exts2 as the job executor. The CUDA kernel for these jobs does not exist yet
    (the product refuses -x splice); this pins the driver half of that row."""
    H, R, api = libs
    rng = np.random.default_rng(11)
    contigs = synth.random_genome(300_000, 29, n_contigs=1, repeat_frac=0.0)
    g = np.frombuffer(bytes(contigs[0]), dtype=np.uint8).copy()
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    reads = []
    for i in range(36):
        pos = int(rng.integers(2000, len(g) - 60000))
        exons = []
        strand_rev = i % 2 == 1
        for k in range(int(rng.integers(2, 6))):
            el = int(rng.integers(90, 400))
            exons.append((pos, pos + el))
            il = int(rng.integers(150, 6000))
            if i % 3 != 2:  # plant canonical splice signals on the transcript strand: GT..AG (or CT..AC on the minus strand)
                d, a_ = (b"GT", b"AG") if not strand_rev else (b"CT", b"AC")
                g[pos + el:pos + el + 2] = list(d); g[pos + el + il - 2:pos + el + il] = list(a_)
            pos += el + il
        tr = np.concatenate([g[s:e] for s, e in exons])
        if strand_rev:
            tr = comp[tr[::-1]]
        reads.append(synth.mutate_ascii(tr, rng, 0.03))
    contigs = [g.tobytes()]
    n, _ = run_case(H, R, api, contigs, ["chr0"], reads, preset="splice", w=5, is_cdna=1)
    assert n >= 30


def _spliced_set(seed, n_reads, glen=300_000):
    """cDNA reads over a random genome; returns (genome bytes, reads, introns [(st, en, strand)])"""
    rng = np.random.default_rng(seed)
    g = np.frombuffer(bytes(synth.random_genome(glen, seed + 1, n_contigs=1, repeat_frac=0.0)[0]), dtype=np.uint8).copy()
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    reads, introns = [], []
    for i in range(n_reads):
        pos = int(rng.integers(2000, len(g) - 60000)); exons = []; rev = i % 2 == 1
        for k in range(int(rng.integers(2, 6))):
            el = int(rng.integers(60, 300)); exons.append((pos, pos + el)); il = int(rng.integers(150, 5000))
            if i % 3 == 0:  # canonical signals for a third of the transcripts only: the others depend on the annotation
                d, a_ = (b"GT", b"AG") if not rev else (b"CT", b"AC")
                g[pos + el:pos + el + 2] = list(d); g[pos + el + il - 2:pos + el + il] = list(a_)
            introns.append((pos + el, pos + el + il, -1 if rev else 1))
            pos += el + il
        introns.pop()  # the stretch after the last exon is not an intron
        tr = np.concatenate([g[s:e] for s, e in exons])
        reads.append(synth.mutate_ascii(comp[tr[::-1]] if rev else tr, rng, 0.04))
    return g.tobytes(), reads, introns


def _write_bed(path, introns, rng):
    """BED6 lines (one intron each) in random order with duplicates, shifted decoys, a strandless line and an unknown contig"""
    lines = []
    for st, en, sd in introns:
        lines.append("chr0\t%d\t%d\tj\t%d\t%s" % (st, en, int(rng.integers(0, 100)), "+" if sd > 0 else "-"))
        if rng.random() < 0.3:
            lines.append(lines[-1])
        if rng.random() < 0.3:
            d = int(rng.integers(-6, 7))
            lines.append("chr0\t%d\t%d\tdecoy\t0\t%s" % (st + d, en + d, "+" if rng.random() < 0.5 else "-"))
    lines.append("chr0\t100\t900\tnostrand\t0\t.")
    lines.append("chrUn\t100\t900\tx\t0\t+")
    # a BED12 transcript: three blocks -> two introns (3000-3400 and 3600-5000)
    lines.append("chr0\t2900\t5100\ttx\t0\t+\t2900\t5100\t0\t3\t100,200,100,\t0,500,2100,")
    order = rng.permutation(len(lines))
    open(path, "w").write("".join(lines[i] + "\n" for i in order))


def test_bed_reader_and_junction_flags_match_reference(libs, tmp_path):
    """mm_idx_bed_read / mm_idx_bed_junc (index.c:672-826) restated in csrc/annot.h vs the reference library: merged interval lists
    and the per-window flag arrays on random windows"""
    H, R, api = libs
    rng = np.random.default_rng(3)
    g, reads, introns = _spliced_set(50, 40)
    bed = str(tmp_path / "anno.bed"); _write_bed(bed, introns, rng)
    mi, keep = build_ref_index(R, [g], ["chr0"], w=5)
    R.mm_idx_bed_read(mi, bed.encode(), 1)
    H.hs_bed_read.restype = C.c_void_p
    I = H.hs_bed_read(mi, bed.encode(), 1)
    n = H.hs_bed_n(C.c_void_p(I), 0)
    assert n >= len(set(introns))
    for _ in range(400):
        st = int(rng.integers(0, len(g) - 10)); en = min(len(g), st + int(rng.integers(1, 30000)))
        a = np.zeros(en - st, dtype=np.uint8); b = np.full(en - st, 77, dtype=np.uint8)
        ra = R.mm_idx_bed_junc(mi, 0, st, en, a.ctypes.data_as(C.c_void_p))
        rb = H.hs_bed_junc(C.c_void_p(I), 1, 0, st, en, b.ctypes.data_as(C.c_void_p))
        assert ra == rb and np.array_equal(a, b), (st, en)
    H.hs_bed_free(C.c_void_p(I), 1)


def test_bed_reader_unstable_merge_order_matches_reference(libs, tmp_path):
    """more than 64 lines per contig take the reference's unstable radix sort (index.c:776-783, ksort.h:101-151): of several lines with the
    same (st, en) but different strands, the one that sort happens to put first names the merged intron's strand. Round 1 used a
    stable sort here and lost a junction bonus on the device suite (same CIGAR, AS off by 2*junc_bonus)."""
    H, R, api = libs
    rng = np.random.default_rng(15)
    g, reads, introns = _spliced_set(70, 150, glen=800_000)
    bed = str(tmp_path / "anno.bed")
    lines = []
    for st, en, sd in introns:
        for _ in range(int(rng.integers(1, 4))):  # the same interval on both strands, several times
            lines.append("chr0\t%d\t%d\tj\t%d\t%s" % (st, en, int(rng.integers(0, 100)), "+-"[int(rng.integers(0, 2))]))
        if rng.random() < 0.5:  # same start, other ends
            lines.append("chr0\t%d\t%d\tk\t0\t%s" % (st, en + int(rng.integers(1, 300)), "+-"[int(rng.integers(0, 2))]))
    order = rng.permutation(len(lines))
    open(bed, "w").write("".join(lines[i] + "\n" for i in order))
    assert len(lines) > 300
    mi, keep = build_ref_index(R, [g], ["chr0"], w=5)
    R.mm_idx_bed_read(mi, bed.encode(), 1)
    H.hs_bed_read.restype = C.c_void_p
    I = H.hs_bed_read(mi, bed.encode(), 1)
    for _ in range(300):
        st = int(rng.integers(0, len(g) - 10)); en = min(len(g), st + int(rng.integers(1, 60000)))
        a = np.zeros(en - st, dtype=np.uint8); b = np.full(en - st, 77, dtype=np.uint8)
        ra = R.mm_idx_bed_junc(mi, 0, st, en, a.ctypes.data_as(C.c_void_p))
        rb = H.hs_bed_junc(C.c_void_p(I), 1, 0, st, en, b.ctypes.data_as(C.c_void_p))
        assert ra == rb and np.array_equal(a, b), (st, en)
    a = np.zeros(len(g), dtype=np.uint8); b = np.zeros(len(g), dtype=np.uint8)
    R.mm_idx_bed_junc(mi, 0, 0, len(g), a.ctypes.data_as(C.c_void_p)); H.hs_bed_junc(C.c_void_p(I), 1, 0, 0, len(g), b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)
    H.hs_bed_free(C.c_void_p(I), 1)


def test_driver_with_junction_annotation(libs, tmp_path):
    """-x splice --junc-bed: the junction flags reach ksw_exts2 exactly as mm_get_junc / mm_idx_bed_junc pass them (per-call windows,
    reversed for the left extension), for annotated non-canonical introns, decoys next to the true sites and both strands"""
    H, R, api = libs
    rng = np.random.default_rng(5)
    g, reads, introns = _spliced_set(60, 36)
    bed = str(tmp_path / "anno.bed"); _write_bed(bed, introns, rng)
    n, _ = run_case(H, R, api, [g], ["chr0"], reads, preset="splice", w=5, is_cdna=1, bed=bed)
    assert n >= 30


def _write_spsc(path, g, introns, rng):
    """splice-score lines (contig, position, strand, D/A, score) at the true sites, at decoys and at random places, with duplicates"""
    lines = []
    for st, en, sd in introns:
        s_ = "+" if sd > 0 else "-"
        d_pos, a_pos = (st, en - 1) if sd > 0 else (en - 1, st)  # first/last intron base in transcript order
        lines.append("chr0\t%d\t%s\tD\t%d" % (d_pos, s_, int(rng.integers(3, 14))))
        lines.append("chr0\t%d\t%s\tA\t%d" % (a_pos, s_, int(rng.integers(3, 14))))
        if rng.random() < 0.3:
            lines.append("chr0\t%d\t%s\tD\t%d" % (d_pos, s_, int(rng.integers(-8, 20))))
    for _ in range(len(introns) * 4):
        lines.append("chr0\t%d\t%s\t%s\t%d" % (int(rng.integers(0, len(g))), "+-"[int(rng.integers(0, 2))], "DA"[int(rng.integers(0, 2))], int(rng.integers(-12, 12))))
    lines += ["chr0\t0\t+\tD\t5", "chrUn\t50\t+\tD\t5", "chr0\t77\t+\tX\t5", "chr0\t88\t+"]
    order = rng.permutation(len(lines))
    open(path, "w").write("".join(lines[i] + "\n" for i in order))


def test_splice_score_reader_matches_reference(libs, tmp_path):
    """mm_idx_spsc_read2 / mm_idx_spsc_get (index.c:963-1075) restated in csrc/annot.h vs the reference library on random windows"""
    H, R, api = libs
    rng = np.random.default_rng(4)
    g, reads, introns = _spliced_set(50, 40)
    fn = str(tmp_path / "sc.txt"); _write_spsc(fn, g, introns, rng)
    mi, keep = build_ref_index(R, [g], ["chr0"], w=5)
    R.mm_idx_spsc_read2.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_float]
    R.mm_idx_spsc_get.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
    R.mm_idx_spsc_get.restype = C.c_int64
    H.hs_spsc_read.restype = C.c_void_p; H.hs_spsc_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float]
    H.hs_spsc_get.restype = C.c_int64; H.hs_spsc_get.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
    for scale in (0.7, 1.0):
        R.mm_idx_spsc_read2(mi, fn.encode(), 30, scale)
        S = H.hs_spsc_read(mi, fn.encode(), 30, scale)
        for _ in range(300):
            st = int(rng.integers(0, len(g) - 10)); en = min(len(g), st + int(rng.integers(1, 20000))); rev = int(rng.integers(0, 2))
            a = np.zeros(en - st, dtype=np.uint8); b = np.zeros(en - st, dtype=np.uint8)
            ra = R.mm_idx_spsc_get(mi, 0, st, en, rev, a.ctypes.data_as(C.c_void_p))
            rb = H.hs_spsc_get(S, mi, 0, st, en, rev, b.ctypes.data_as(C.c_void_p))
            assert ra == rb and np.array_equal(a, b), (st, en, rev)
        H.hs_spsc_free(C.c_void_p(S), 1)


def test_driver_with_splice_scores(libs, tmp_path):
    """-x splice --spsc: KSW_EZ_SPLICE_SCORE on every spliced call (align.c:688) and junc[] bytes from mm_idx_spsc_get on the strand
    SPLICE_REV selects (align.c:638-640), reversed for the left extension"""
    H, R, api = libs
    rng = np.random.default_rng(6)
    g, reads, introns = _spliced_set(61, 36)
    fn = str(tmp_path / "sc.txt"); _write_spsc(fn, g, introns, rng)
    n, _ = run_case(H, R, api, [g], ["chr0"], reads, preset="splice", w=5, is_cdna=1, spsc=fn)
    assert n >= 30


def test_driver_in_query_strand_mode(libs):
    """--qstrand (MM_F_QSTRAND | MM_F_NO_INV, main.c:252): the query stays forward and a reverse hit is aligned against the other
    strand of the target (mm_idx_getseq2, align.c:780-783,815-818,875-878,899-901), coordinates on that strand; PAF prints them
    flipped (format.c:440-443) and cs/MD use the same view"""
    H, R, api = libs
    contigs = synth.random_genome(200_000, 27, n_contigs=2, repeat_frac=0.1)
    reads = synth.make_reads(contigs, 40, 3000, 0.08, 127, chimeric_frac=0.1)

    def qstrand(mo):
        mo.flag |= 0x100000000 | 0x200000000  # MM_F_QSTRAND | MM_F_NO_INV
    n, _ = run_case(H, R, api, contigs, ["chr0", "chr1"], reads, tweak=qstrand)
    assert n >= 30


def test_driver_on_hifi_and_single_affine(libs):
    """other scoring regimes through the same driver: map-hifi (k19/w19, a1 b4 q6,26) and a single-affine setting (q == q2:
    the reference switches to ksw_extz2 there, align.c:360; the product keeps the dual-affine jobs with equal terms)"""
    H, R, api = libs
    contigs = synth.random_genome(250_000, 37, n_contigs=2, repeat_frac=0.1)
    reads = synth.make_reads(contigs, 25, 6000, 0.01, 137, chimeric_frac=0.1)
    n, _ = run_case(H, R, api, contigs, ["chr0", "chr1"], reads, preset="map-hifi", w=19, k=19)
    assert n >= 20

    def single_affine(mo):
        mo.q = mo.q2 = 6; mo.e = mo.e2 = 2
    reads = synth.make_reads(contigs, 25, 3000, 0.10, 138, chimeric_frac=0.1)
    n, _ = run_case(H, R, api, contigs, ["chr0", "chr1"], reads, preset="map-ont", tweak=single_affine)
    assert n >= 20
