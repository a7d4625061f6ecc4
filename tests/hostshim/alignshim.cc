// tests/hostshim/alignshim.cc -- TEST INFRASTRUCTURE: runs the product's replayable alignment driver (minimap2_b200/csrc/align.cc,
// the restatement of mm_align_skeleton / mm_align1 / mm_align1_inv, align.c:645-1120) on the CPU, with the oracle's ksw2
// (oracle/mm2o_ksw2.c) standing in for the GPU as the executor of the ksw jobs the driver requests -- the same wave/replay
// protocol as csrc/map.cu. tests/test_aligndriver_vs_ref.py compares the result with the reference's mm_align_skeleton on the
// same chains, so the driver logic is checked without a GPU and independently of the CUDA kernels.
#include "hostlogic.h"
#include "../../oracle/mm2o.h"
#include "annot.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
extern int mm_verbose;

// mm_idx_getseq on the public part of mm_idx_t (index.c:164-175): the product's own lives in csrc/index.cu (CUDA file)
int mm_idx_getseq(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{
	if (rid >= mi->n_seq || st >= mi->seq[rid].len) return -1;
	if (en > mi->seq[rid].len) en = mi->seq[rid].len;
	const uint64_t st1 = mi->seq[rid].offset + st, en1 = mi->seq[rid].offset + en;
	for (uint64_t i = st1; i < en1; ++i) seq[i - st1] = (uint8_t)mmx_seq4_get(mi->S, i);
	return (int)(en - st);
}

static uint8_t nt4_of(uint8_t c) { return mmx_nt4(c); }

// Executes one requested job with the oracle; the CIGAR is kept alive in `store`.
static KswDone run_job(const mm_mapopt_t *opt, const mm_idx_t *mi, const int8_t *mat, const uint8_t *q_fwd, int64_t q_dev_off, const mmb_ksw_job_t &j,
					   std::vector<uint32_t*> &store)
{
	std::vector<uint8_t> q(j.qlen > 0? j.qlen : 0), t(j.tlen > 0? j.tlen : 0);
	for (int i = 0; i < j.qlen; ++i) {
		uint8_t c = q_fwd[j.q_start - q_dev_off + (int64_t)i * j.q_step];
		if ((j.flag & MMB_JOB_Q_COMP) && c < 4) c = 3 - c;
		q[i] = c;
	}
	for (int i = 0; i < j.tlen; ++i) {
		uint8_t c = (uint8_t)mmx_seq4_get(mi->S, (uint64_t)(j.t_start + (int64_t)i * j.t_step));
		if ((j.flag & MMB_JOB_T_COMP) && c < 4) c = 3 - c;
		t[i] = c;
	}
	if (const char *dump = getenv("HS_DUMP_JOBS")) { // debugging aid: the jobs with their sequences, one text record each
		FILE *fp = fopen(dump, "a");
		fprintf(fp, "J %d %d %d %d %d %d\n", j.qlen, j.tlen, j.w, j.zdrop, j.end_bonus, j.flag);
		for (int i = 0; i < j.qlen; ++i) fputc('0' + q[i], fp);
		fputc('\n', fp);
		for (int i = 0; i < j.tlen; ++i) fputc('0' + t[i], fp);
		fputc('\n', fp);
		fclose(fp);
	}
	KswDone d;
	memset(&d, 0, sizeof(d));
	d.r.zd_max = -1; // the scan of mm_test_zdrop is left to the driver (host path)
	if (j.flag & MMB_JOB_LL) {
		int qe = -1, te = -1;
		d.r.score = mm2o_ll_i16(j.qlen, q.data(), j.tlen, t.data(), 5, mat, j.w, j.zdrop, &qe, &te);
		d.r.max_q = qe, d.r.max_t = te;
		d.cig = nullptr;
		return d;
	}
	mm2o_ez_t ez;
	memset(&ez, 0, sizeof(ez));
	if (j.flag & MMB_JOB_SPLICE) { // ksw_exts2 as mm_align_pair calls it (align.c:352-355), junc[] as mm_get_junc fills it (align.c:638-643,787-790)
		std::vector<uint8_t> junc;
		if ((mi->I || mi->spsc) && j.tlen > 0) { // mm_get_junc (align.c:638-643): splice scores take precedence over the BED annotation
			const int64_t lo = j.t_step > 0? j.t_start : j.t_start - (j.tlen - 1);
			int32_t rid = 0;
			while (rid + 1 < (int32_t)mi->n_seq && (int64_t)mi->seq[rid + 1].offset <= lo) ++rid;
			const int32_t st = (int32_t)(lo - (int64_t)mi->seq[rid].offset);
			junc.resize(j.tlen);
			if (mi->spsc) mmx_spsc_get((const mm_idx_spsc_s*)mi->spsc, (int32_t)mi->n_seq, rid, mi->seq[rid].len, st, st + j.tlen, !!(j.flag & MMB_KSW_SPLICE_REV), junc.data());
			else mmx_bed_junc((const mm_idx_intv_s*)mi->I, (int32_t)mi->n_seq, rid, st, st + j.tlen, junc.data());
			if (j.t_step < 0) std::reverse(junc.begin(), junc.end());
		}
		mm2o_exts2(j.qlen, q.data(), j.tlen, t.data(), 5, mat, opt->q, opt->e, opt->q2, opt->noncan, j.zdrop, j.end_bonus, opt->junc_bonus, opt->junc_pen,
				   j.flag & 0x1fff, junc.empty()? nullptr : junc.data(), &ez);
	}
	else mm2o_extd2(j.qlen, q.data(), j.tlen, t.data(), 5, mat, opt->q, opt->e, opt->q2, opt->e2, j.w, j.zdrop, j.end_bonus, j.flag & 0xff, &ez);
	d.r.max = ez.max, d.r.zdropped = ez.zdropped, d.r.max_q = ez.max_q, d.r.max_t = ez.max_t, d.r.mqe = ez.mqe, d.r.mqe_t = ez.mqe_t;
	d.r.mte = ez.mte, d.r.mte_q = ez.mte_q, d.r.score = ez.score, d.r.n_cigar = ez.n_cigar, d.r.reach_end = ez.reach_end;
	d.cig = ez.cigar;
	store.push_back(ez.cigar);
	return d;
}

// host-cost profiling (tests/hostshim/profile_driver.py): cycles spent in requesting replays [0] and in the final replay [1]
static uint64_t g_replay_cycles[2];
void hs_prof_enable(int on) { g_hp_on = on != 0; g_replay_cycles[0] = g_replay_cycles[1] = 0; memset(g_hp, 0, sizeof(g_hp)); }
void hs_prof_read(uint64_t *sections, uint64_t *replays)
{
	hl_hp_flush();
	memcpy(sections, g_hp, sizeof(g_hp));
	replays[0] = g_replay_cycles[0], replays[1] = g_replay_cycles[1];
}
int hs_prof_n_sections(void) { return HP_N; }

// regs: libc-malloc'd array of n_regs hits (after chain post-processing), a: their anchors. Returns the aligned hits.
mm_reg1_t *hs_align_skeleton(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, const char *qstr, int *n_regs_, const mm_reg1_t *regs0, int n_a, const m128 *a0,
							 int *n_waves)
{
	int8_t mat[25];
	{ // align.c:11-38 for m = 5
		const int8_t a = (int8_t)(opt->a < 0? -opt->a : opt->a), b = (int8_t)(opt->b > 0? -opt->b : opt->b), sa = (int8_t)(opt->sc_ambi > 0? -opt->sc_ambi : opt->sc_ambi);
		for (int i = 0; i < 4; ++i) { for (int k = 0; k < 4; ++k) mat[i * 5 + k] = i == k? a : b; mat[i * 5 + 4] = sa; }
		for (int k = 0; k < 5; ++k) mat[20 + k] = sa;
		if (!(opt->transition == 0 || opt->transition == opt->b)) {
			const int8_t t = (int8_t)(opt->transition > 0? -opt->transition : opt->transition);
			mat[0 * 5 + 2] = t, mat[1 * 5 + 3] = t, mat[2 * 5 + 0] = t, mat[3 * 5 + 1] = t;
		}
	}
	std::vector<uint8_t> qf(qlen + 16), qr(qlen + 16); // the driver may read 16 bytes past a sequence (block compares)
	for (int i = 0; i < qlen; ++i) { const uint8_t c = nt4_of((uint8_t)qstr[i]); qf[i] = c, qr[qlen - 1 - i] = c < 4? 3 - c : 4; }
	ReadAlign ra;
	ra.reset();
	ra.qlen = qlen, ra.q_dev_off = 1000, ra.qseq[0] = qf.data(), ra.qseq[1] = qr.data(); // any offset: jobs address the query through it
	std::vector<uint32_t*> store;
	std::vector<m128> a(n_a);
	mm_reg1_t *out = nullptr;
	int n_out = 0;
	for (int wave = 0;; ++wave) {
		if (wave > 100000) abort();
		ra.want.clear(); ra.want_slot.clear();
		ra.incomplete = false;
		memcpy(a.data(), a0, sizeof(m128) * n_a);
		int n_regs = *n_regs_;
		mm_reg1_t *regs = (mm_reg1_t*)malloc(sizeof(mm_reg1_t) * (n_regs > 0? n_regs : 1));
		memcpy(regs, regs0, sizeof(mm_reg1_t) * n_regs);
		const uint64_t c0 = __builtin_ia32_rdtsc();
		regs = hl_align_skeleton(opt, mi, ra, &n_regs, regs, n_a, a.data());
		g_replay_cycles[ra.incomplete? 0 : 1] += __builtin_ia32_rdtsc() - c0;
		if (!ra.incomplete) { out = regs, n_out = n_regs; if (n_waves) *n_waves = wave; break; }
		for (int i = 0; i < n_regs; ++i) free(regs[i].p);
		free(regs);
		if (ra.want.empty()) abort(); // no progress
		for (size_t i = 0; i < ra.want.size(); ++i) {
			KswDone d = run_job(opt, mi, mat, qf.data(), ra.q_dev_off, ra.want[i], store);
			ra.done_idx[ra.want_slot[i]] = (int)ra.done.size();
			ra.done.push_back(d);
		}
	}
	for (uint32_t *c : store) free(c);
	*n_regs_ = n_out;
	return out;
}
void hs_free_regs(int n, mm_reg1_t *r) { for (int i = 0; i < n; ++i) free(r[i].p); free(r); }
}

// ---- formatting (csrc/format.cc) next to the reference's format.c ----
extern "C" {
// writes the PAF line of regs[j] (or the SAM record) into buf (capacity cap); returns the length
int hs_write_paf(char *buf, int cap, const mm_idx_t *mi, const char *qname, const char *seq, int qlen, const mm_reg1_t *r, int64_t opt_flag, int rep_len)
{
	std::string s;
	hl_set_seq_for_tags(seq);
	hl_write_paf(s, mi, qname, qlen, r, opt_flag, rep_len);
	if ((int)s.size() >= cap) return -1;
	memcpy(buf, s.data(), s.size()); buf[s.size()] = 0;
	return (int)s.size();
}
int hs_write_sam(char *buf, int cap, const mm_idx_t *mi, const char *qname, const char *seq, const char *qual, int qlen, int reg_idx, int n_regs,
				 const mm_reg1_t *regs, int64_t opt_flag, int rep_len)
{
	std::string s;
	hl_write_sam(s, mi, qname, seq, qual, qlen, reg_idx, n_regs, regs, opt_flag, rep_len);
	if ((int)s.size() >= cap) return -1;
	memcpy(buf, s.data(), s.size()); buf[s.size()] = 0;
	return (int)s.size();
}
}

// ---- junction annotation (csrc/annot.h) next to the reference's mm_idx_bed_read / mm_idx_bed_junc ----
#include <unordered_map>
extern "C" {
// reads fn with the product's BED reader against the contig names of mi (any mm_idx_t); returns the per-contig interval array
void *hs_bed_read(const mm_idx_t *mi, const char *fn, int read_junc)
{
	std::unordered_map<std::string, int> ids;
	for (uint32_t i = 0; i < mi->n_seq; ++i) ids[mi->seq[i].name] = (int)i;
	return mmx_bed_read(fn, mi->n_seq, read_junc, -1, [&](const char *name) { auto it = ids.find(name); return it == ids.end()? -1 : it->second; }, nullptr, nullptr);
}
int hs_bed_n(const void *I, int ctg) { return ((const mm_idx_intv_s*)I)[ctg].n; }
void hs_bed_get(const void *I, int ctg, int k, int32_t *out) { const mm_idx_intv1_t &t = ((const mm_idx_intv_s*)I)[ctg].a[k]; out[0] = t.st, out[1] = t.en, out[2] = t.cnt, out[3] = t.strand; }
int hs_bed_junc(const void *I, int n_seq, int ctg, int st, int en, uint8_t *s) { return mmx_bed_junc((const mm_idx_intv_s*)I, n_seq, ctg, st, en, s); }
void hs_bed_free(void *I, int n_seq) { for (int i = 0; i < n_seq; ++i) free(((mm_idx_intv_s*)I)[i].a); free(I); }
void *hs_spsc_read(const mm_idx_t *mi, const char *fn, int max_sc, float scale)
{
	std::unordered_map<std::string, int> ids;
	for (uint32_t i = 0; i < mi->n_seq; ++i) ids[mi->seq[i].name] = (int)i;
	return mmx_spsc_read(fn, mi->n_seq, max_sc, scale, [&](const char *name) { auto it = ids.find(name); return it == ids.end()? -1 : it->second; },
						 [&](int cid) { return mi->seq[cid].len; }, nullptr);
}
int64_t hs_spsc_get(const void *S, const mm_idx_t *mi, int cid, int64_t st, int64_t en, int rev, uint8_t *sc)
{
	return mmx_spsc_get((const mm_idx_spsc_s*)S, (int32_t)mi->n_seq, cid, mi->seq[cid].len, st, en, rev, sc);
}
void hs_spsc_free(void *S, int n_seq) { for (int i = 0; i < n_seq * 2; ++i) free(((mm_idx_spsc_s*)S)[i].a); free(S); }
}
