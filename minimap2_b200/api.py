"""Python mirror of the reference's mappy binding (python/mappy.pyx: Aligner, Alignment, ThreadBuffer, revcomp) over
the minimap.h C API exported by libminimap2_b200.so, plus a batch entry point (map_batch) that feeds the GPU scheduler
thousands of reads per call. Struct layouts mirror include/minimap.h (asserted against the reference in tests/test_abi.py)."""
import ctypes as C
import numpy as np
from ._lib import lib


class IdxOpt(C.Structure):  # mm_idxopt_t
    _fields_ = [("k", C.c_short), ("w", C.c_short), ("flag", C.c_short), ("bucket_bits", C.c_short),
                ("mini_batch_size", C.c_int64), ("batch_size", C.c_uint64)]


class MapOpt(C.Structure):  # mm_mapopt_t
    _fields_ = [("flag", C.c_int64), ("seed", C.c_int), ("sdust_thres", C.c_int), ("max_qlen", C.c_int),
                ("bw", C.c_int), ("bw_long", C.c_int), ("max_gap", C.c_int), ("max_gap_ref", C.c_int),
                ("max_frag_len", C.c_int), ("max_chain_skip", C.c_int), ("max_chain_iter", C.c_int),
                ("min_cnt", C.c_int), ("min_chain_score", C.c_int), ("chain_gap_scale", C.c_float),
                ("chain_skip_scale", C.c_float), ("rmq_size_cap", C.c_int), ("rmq_inner_dist", C.c_int),
                ("rmq_rescue_size", C.c_int), ("rmq_rescue_ratio", C.c_float), ("mask_level", C.c_float),
                ("mask_len", C.c_int), ("pri_ratio", C.c_float), ("best_n", C.c_int), ("alt_drop", C.c_float),
                ("a", C.c_int), ("b", C.c_int), ("q", C.c_int), ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int),
                ("transition", C.c_int), ("sc_ambi", C.c_int), ("noncan", C.c_int), ("junc_bonus", C.c_int),
                ("junc_pen", C.c_int), ("zdrop", C.c_int), ("zdrop_inv", C.c_int), ("end_bonus", C.c_int),
                ("min_dp_max", C.c_int), ("min_ksw_len", C.c_int), ("anchor_ext_len", C.c_int),
                ("anchor_ext_shift", C.c_int), ("max_clip_ratio", C.c_float), ("rank_min_len", C.c_int),
                ("rank_frac", C.c_float), ("pe_ori", C.c_int), ("pe_bonus", C.c_int), ("jump_min_match", C.c_int32),
                ("mid_occ_frac", C.c_float), ("q_occ_frac", C.c_float), ("min_mid_occ", C.c_int32),
                ("max_mid_occ", C.c_int32), ("mid_occ", C.c_int32), ("max_occ", C.c_int32), ("max_max_occ", C.c_int32),
                ("occ_dist", C.c_int32), ("mini_batch_size", C.c_int64), ("max_sw_mat", C.c_int64),
                ("cap_kalloc", C.c_int64), ("split_prefix", C.c_char_p)]


class IdxSeq(C.Structure):  # mm_idx_seq_t
    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint64), ("len", C.c_uint32), ("is_alt", C.c_uint32)]


class Idx(C.Structure):  # mm_idx_t
    _fields_ = [("b", C.c_int32), ("w", C.c_int32), ("k", C.c_int32), ("flag", C.c_int32), ("n_seq", C.c_uint32),
                ("index", C.c_int32), ("n_alt", C.c_int32), ("seq", C.POINTER(IdxSeq)), ("S", C.POINTER(C.c_uint32)),
                ("B", C.c_void_p), ("I", C.c_void_p), ("spsc", C.c_void_p), ("J", C.c_void_p), ("km", C.c_void_p), ("h", C.c_void_p)]


class Extra(C.Structure):  # mm_extra_t (header; cigar[] follows)
    _fields_ = [("capacity", C.c_uint32), ("dp_score", C.c_int32), ("dp_max", C.c_int32), ("dp_max2", C.c_int32),
                ("dp_max0", C.c_int32), ("n_ambi_ts", C.c_uint32), ("n_cigar", C.c_uint32)]


class Reg1(C.Structure):  # mm_reg1_t
    _fields_ = [("id", C.c_int32), ("cnt", C.c_int32), ("rid", C.c_int32), ("score", C.c_int32),
                ("qs", C.c_int32), ("qe", C.c_int32), ("rs", C.c_int32), ("re", C.c_int32),
                ("parent", C.c_int32), ("subsc", C.c_int32), ("as_", C.c_int32), ("mlen", C.c_int32), ("blen", C.c_int32),
                ("n_sub", C.c_int32), ("score0", C.c_int32), ("bits", C.c_uint32), ("hash", C.c_uint32), ("div", C.c_float),
                ("p", C.POINTER(Extra))]

    @property
    def mapq(self): return self.bits & 0xff
    @property
    def split(self): return self.bits >> 8 & 3
    @property
    def rev(self): return self.bits >> 10 & 1
    @property
    def inv(self): return self.bits >> 11 & 1
    @property
    def sam_pri(self): return self.bits >> 12 & 1


MM_F_CIGAR, MM_F_OUT_CG = 0x004, 0x020
_setup_done = False


def _setup():
    global _setup_done
    L = lib()
    if _setup_done:
        return L
    L.mm_set_opt.restype = C.c_int
    L.mm_set_opt.argtypes = [C.c_char_p, C.POINTER(IdxOpt), C.POINTER(MapOpt)]
    L.mm_check_opt.restype = C.c_int
    L.mm_mapopt_update.argtypes = [C.POINTER(MapOpt), C.POINTER(Idx)]
    L.mm_idx_reader_open.restype = C.c_void_p
    L.mm_idx_reader_open.argtypes = [C.c_char_p, C.POINTER(IdxOpt), C.c_char_p]
    L.mm_idx_reader_read.restype = C.POINTER(Idx)
    L.mm_idx_reader_read.argtypes = [C.c_void_p, C.c_int]
    L.mm_idx_reader_close.argtypes = [C.c_void_p]
    L.mm_idx_destroy.argtypes = [C.POINTER(Idx)]
    L.mm_idx_index_name.argtypes = [C.POINTER(Idx)]
    L.mm_idx_str.restype = C.POINTER(Idx)
    L.mm_idx_str.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.mm_map_batch.restype = C.c_int
    L.mm_map_batch.argtypes = [C.POINTER(Idx), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.POINTER(MapOpt), C.c_int]
    L.mm_map.restype = C.POINTER(Reg1)
    L.mm_map.argtypes = [C.POINTER(Idx), C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.POINTER(MapOpt), C.c_char_p]
    L.mm_tbuf_init.restype = C.c_void_p
    L.mm_tbuf_destroy.argtypes = [C.c_void_p]
    L.mmb_synth_index.restype = C.POINTER(Idx)
    L.mmb_synth_index.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.mmb_synth_reads.restype = C.c_int
    L.mmb_synth_reads.argtypes = [C.POINTER(Idx), C.c_int, C.c_int, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_void_p]
    L.mmb_idx_write_fasta.restype = C.c_int
    L.mmb_idx_write_fasta.argtypes = [C.POINTER(Idx), C.c_char_p]
    L.mmb_default_ctx_c.restype = C.c_void_p
    L.mmb_free.argtypes = [C.c_void_p]
    L.mmb_profile_bytes.restype = C.c_uint64
    L.mmb_profile_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mmb_profile_scopes.restype = C.c_uint64
    L.mmb_profile_scopes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mmb_set_resident_reads.argtypes = [C.c_int]
    L.mmb_set_groups.argtypes = [C.c_int]
    L.mmb_profile_enable_all.argtypes = [C.c_int]
    L.mmb_profile_ms_all.restype = C.c_double
    L.mmb_profile_ms_all.argtypes = [C.c_int, C.c_int]
    for nm in ("mmb_profile_units_all", "mmb_profile_bytes_all", "mmb_profile_scopes_all"):
        getattr(L, nm).restype = C.c_uint64
        getattr(L, nm).argtypes = [C.c_int, C.c_int]
    L.mmb_launch_count_all.restype = C.c_uint64
    L.mmb_launch_count_all.argtypes = [C.c_int]
    L.mmb_last_d2h_bytes.restype = C.c_uint64
    _setup_done = True
    return L


class Alignment:
    """mappy.Alignment equivalent (python/mappy.pyx:25-105)"""

    def __init__(self, ctg, cl, cs, ce, strand, qs, qe, mapq, cigar, is_primary, mlen, blen, NM, trans_strand, seg_id):
        self.ctg, self.ctg_len, self.r_st, self.r_en = ctg, cl, cs, ce
        self.strand, self.q_st, self.q_en, self.mapq = strand, qs, qe, mapq
        self.cigar, self.is_primary, self.mlen, self.blen, self.NM = cigar, is_primary, mlen, blen, NM
        self.trans_strand, self.read_num = trans_strand, seg_id + 1

    @property
    def cigar_str(self):
        return "".join("%d%s" % (c[0], "MIDNSHP=XB"[c[1]]) for c in self.cigar)

    def __str__(self):
        strand = "+" if self.strand > 0 else "-" if self.strand < 0 else "?"
        tp = "tp:A:P" if self.is_primary else "tp:A:S"
        ts = "ts:A:." if self.trans_strand == 0 else "ts:A:+" if self.trans_strand > 0 else "ts:A:-"
        return "\t".join(map(str, [self.q_st, self.q_en, strand, self.ctg, self.ctg_len, self.r_st, self.r_en, self.mlen,
                                   self.blen, self.mapq, tp, ts, "cg:Z:" + self.cigar_str]))


def _reg_to_alignment(mi, r):
    seq = mi.contents.seq[r.rid]
    cigar, NM, ts = [], 0, 0
    if r.p:
        ex = r.p.contents
        base = C.addressof(ex) + C.sizeof(Extra)
        arr = (C.c_uint32 * ex.n_cigar).from_address(base)
        cigar = [[c >> 4, c & 0xf] for c in arr]
        n_ambi = ex.n_ambi_ts & 0x3fffffff
        t = ex.n_ambi_ts >> 30
        NM = r.blen - r.mlen + n_ambi
        ts = 1 if t == 1 else -1 if t == 2 else 0
    return Alignment(seq.name.decode(), seq.len, r.rs, r.re, -1 if r.rev else 1, r.qs, r.qe, r.mapq, cigar, r.id == r.parent,
                     r.mlen, r.blen, NM, ts, r.bits >> 16 & 0xff)


class Aligner:
    """mappy.Aligner equivalent (python/mappy.pyx:116-254) backed by the GPU batch scheduler."""

    def __init__(self, fn_idx_in=None, preset=None, k=None, w=None, min_cnt=None, min_chain_score=None, min_dp_score=None,
                 bw=None, bw_long=None, best_n=None, n_threads=3, fn_idx_out=None, extra_flags=None, seq=None, scoring=None,
                 _idx=None):
        L = _setup()
        self.idx_opt, self.map_opt = IdxOpt(), MapOpt()
        L.mm_set_opt(None, C.byref(self.idx_opt), C.byref(self.map_opt))
        if preset is not None:
            if L.mm_set_opt(preset.encode(), C.byref(self.idx_opt), C.byref(self.map_opt)) != 0:
                raise ValueError("unknown preset %r" % preset)
        self.map_opt.flag |= MM_F_CIGAR
        self.idx_opt.batch_size = 0x7fffffffffffffff
        if k is not None: self.idx_opt.k = k
        if w is not None: self.idx_opt.w = w
        if min_cnt is not None: self.map_opt.min_cnt = min_cnt
        if min_chain_score is not None: self.map_opt.min_chain_score = min_chain_score
        if min_dp_score is not None: self.map_opt.min_dp_max = min_dp_score
        if bw is not None: self.map_opt.bw = bw
        if bw_long is not None: self.map_opt.bw_long = bw_long
        if best_n is not None: self.map_opt.best_n = best_n
        if extra_flags is not None: self.map_opt.flag |= extra_flags
        if scoring is not None and len(scoring) >= 4:
            self.map_opt.a, self.map_opt.b, self.map_opt.q, self.map_opt.e = scoring[:4]
            self.map_opt.q2, self.map_opt.e2 = self.map_opt.q, self.map_opt.e
            if len(scoring) >= 6:
                self.map_opt.q2, self.map_opt.e2 = scoring[4], scoring[5]
        self.n_threads = n_threads
        self._idx = None
        if _idx is not None:
            self._idx = _idx
            L.mm_mapopt_update(C.byref(self.map_opt), self._idx)
        elif seq is None:
            r = L.mm_idx_reader_open(fn_idx_in.encode(), C.byref(self.idx_opt), fn_idx_out.encode() if fn_idx_out else None)
            if r:
                self._idx = L.mm_idx_reader_read(r, n_threads)
                L.mm_idx_reader_close(r)
                if self._idx:
                    L.mm_mapopt_update(C.byref(self.map_opt), self._idx)
                    L.mm_idx_index_name(self._idx)
        else:
            seqs = (C.c_char_p * 1)(seq.encode() if isinstance(seq, str) else seq)
            names = (C.c_char_p * 1)(b"N/A")
            self._idx = L.mm_idx_str(self.idx_opt.w, self.idx_opt.k, self.idx_opt.flag & 1, self.idx_opt.bucket_bits, 1, seqs, names)
            L.mm_mapopt_update(C.byref(self.map_opt), self._idx)
            self.map_opt.mid_occ = 1000

    def __bool__(self):
        return bool(self._idx)

    def close(self):
        if self._idx:
            lib().mm_idx_destroy(self._idx)
            self._idx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prepare_batch(self, buf, qlens, names=None):
        """Pre-builds the pointer arrays mm_map_batch() takes (so repeated calls over the same buffers cost nothing in Python)."""
        n = len(qlens)
        qlens = np.ascontiguousarray(qlens, dtype=np.int32)
        offs = np.zeros(n, dtype=np.uint64)
        if n > 1:
            offs[1:] = np.cumsum(qlens[:-1].astype(np.uint64))
        seq_ptrs = (offs + np.uint64(buf.ctypes.data)).astype(np.uint64)
        name_arr = None
        if names is not None:
            keep = [s if isinstance(s, bytes) else s.encode() for s in names]
            name_arr = ((C.c_char_p * n)(*keep), keep)
        return dict(n=n, qlens=qlens, seq_ptrs=seq_ptrs, names=name_arr, buf=buf)

    def map_prepared(self, b, n_threads=None):
        L = _setup()
        n = b["n"]
        n_regs = np.zeros(n, dtype=np.int32)
        rep_len = np.zeros(n, dtype=np.int32)
        regs = np.zeros(n, dtype=np.uint64)
        name_ptrs = C.cast(b["names"][0], C.c_void_p) if b["names"] is not None else None
        L.mm_map_batch(self._idx, n, b["qlens"].ctypes.data, b["seq_ptrs"].ctypes.data, name_ptrs, n_regs.ctypes.data, regs.ctypes.data,
                       rep_len.ctypes.data, C.byref(self.map_opt), n_threads or self.n_threads)
        return n_regs, regs, rep_len

    def map_batch_raw(self, buf, qlens, names=None, n_threads=None):
        """buf: contiguous uint8 array holding the reads back to back (ASCII); qlens: int32 array.
        Returns (n_regs int32[n], regs pointer array, rep_len int32[n]); call free_batch() on the result."""
        return self.map_prepared(self.prepare_batch(buf, qlens, names), n_threads)

    @staticmethod
    def free_batch(n_regs, regs):
        L = _setup()
        for i in range(len(n_regs)):
            if regs[i]:
                arr = C.cast(C.c_void_p(int(regs[i])), C.POINTER(Reg1))
                for j in range(n_regs[i]):
                    if arr[j].p:
                        L.mmb_free(C.cast(arr[j].p, C.c_void_p))
                L.mmb_free(C.c_void_p(int(regs[i])))

    def map(self, seq, name=None):
        """generator of Alignment objects for one read (mappy.Aligner.map)"""
        L = _setup()
        s = seq if isinstance(seq, bytes) else seq.encode()
        n = C.c_int(0)
        regs = L.mm_map(self._idx, len(s), s, C.byref(n), None, C.byref(self.map_opt), name.encode() if isinstance(name, str) else name)
        try:
            for i in range(n.value):
                yield _reg_to_alignment(self._idx, regs[i])
        finally:
            for i in range(n.value):
                if regs[i].p:
                    L.mmb_free(C.cast(regs[i].p, C.c_void_p))
            if n.value:
                L.mmb_free(C.cast(regs, C.c_void_p))

    @property
    def k(self): return self._idx.contents.k
    @property
    def w(self): return self._idx.contents.w
    @property
    def n_seq(self): return self._idx.contents.n_seq
    @property
    def seq_names(self):
        return [self._idx.contents.seq[i].name.decode() for i in range(self.n_seq)]


def revcomp(seq):
    tab = bytes.maketrans(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")
    s = seq if isinstance(seq, bytes) else seq.encode()
    r = s.translate(tab)[::-1]
    return r if isinstance(seq, bytes) else r.decode()
