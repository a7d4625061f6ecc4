// minimap2_b200/csrc/hits.cc -- host post-processing of chains/hits (cheap, ~1% of the reference's time).
// Re-implements the decision logic of reference hit.c (mm_gen_regs :52, mm_split_reg :106, mm_set_parent :125,
// mm_hit_sort :188, mm_set_sam_pri :225, mm_sync_regs :236, mm_select_sub :255, mm_filter_strand_retained :283,
// mm_filter_regs :301, mm_squeeze_a :322, mm_set_mapq2 :432) and esterr.c (mm_est_err :30) on std::vector scratch.
// Sort tie orders use the exact radix-sort emulation of mm_algo.cuh (mmx_rs_sort), as the reference's are unstable.
#include "hostlogic.h"
#include <cmath>
#include <cstring>
#include <cstdlib>

static void sort128(std::vector<m128> &v)
{
	std::vector<int32_t> stk((size_t)mmx_rs_stack_len((int64_t)v.size()));
	mmx_rs_sort(v.data(), (int64_t)v.size(), stk.data(), KeyX128());
}
static void sort64(std::vector<uint64_t> &v)
{
	std::vector<int32_t> stk((size_t)mmx_rs_stack_len((int64_t)v.size()));
	mmx_rs_sort(v.data(), (int64_t)v.size(), stk.data(), KeyU64());
}

static inline void cal_fuzzy_len(mm_reg1_t *r, const m128 *a) // hit.c:8-22
{
	r->mlen = r->blen = 0;
	if (r->cnt <= 0) return;
	r->mlen = r->blen = (int32_t)(a[r->as].y >> 32 & 0xff);
	for (int i = r->as + 1; i < r->as + r->cnt; ++i) {
		int span = (int)(a[i].y >> 32 & 0xff);
		int tl = (int32_t)a[i].x - (int32_t)a[i-1].x;
		int ql = (int32_t)a[i].y - (int32_t)a[i-1].y;
		r->blen += tl > ql? tl : ql;
		r->mlen += tl > span && ql > span? span : tl < ql? tl : ql;
	}
}

void hl_reg_set_coor(mm_reg1_t *r, int32_t qlen, const m128 *a, int is_qstrand) // hit.c:24-39
{
	int32_t k = r->as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
	r->rev = a[k].x >> 63;
	r->rid = (int32_t)(a[k].x << 1 >> 33);
	r->rs = (int32_t)a[k].x + 1 > q_span? (int32_t)a[k].x + 1 - q_span : 0;
	r->re = (int32_t)a[k + r->cnt - 1].x + 1;
	if (!r->rev || is_qstrand) {
		r->qs = (int32_t)a[k].y + 1 - q_span;
		r->qe = (int32_t)a[k + r->cnt - 1].y + 1;
	} else {
		r->qs = qlen - ((int32_t)a[k + r->cnt - 1].y + 1);
		r->qe = qlen - ((int32_t)a[k].y + 1 - q_span);
	}
	cal_fuzzy_len(r, a);
}

// chains -> hits, best score first with hash-mixed tie breaking (hit.c:52-88). Returns a libc-calloc'd array.
mm_reg1_t *hl_gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const m128 *a, int is_qstrand)
{
	if (n_u <= 0) return 0;
	std::vector<m128> z(n_u);
	int k = 0;
	for (int i = 0; i < n_u; ++i) {
		uint32_t h = (uint32_t)mmx_hash64_full((mmx_hash64_full(a[k].x) + mmx_hash64_full(a[k].y)) ^ hash);
		z[i].x = u[i] ^ h;
		z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
		k += (int32_t)u[i];
	}
	sort128(z);
	for (int i = 0; i < n_u >> 1; ++i) std::swap(z[i], z[n_u - 1 - i]);
	mm_reg1_t *r = (mm_reg1_t*)calloc(n_u, sizeof(mm_reg1_t));
	for (int i = 0; i < n_u; ++i) {
		mm_reg1_t *ri = &r[i];
		ri->id = i;
		ri->parent = HL_PARENT_UNSET;
		ri->score = ri->score0 = (int32_t)(z[i].x >> 32);
		ri->hash = (uint32_t)z[i].x;
		ri->cnt = (int32_t)z[i].y;
		ri->as = (int32_t)(z[i].y >> 32);
		ri->div = -1.0f;
		hl_reg_set_coor(ri, qlen, a, is_qstrand);
	}
	return r;
}

static inline int alt_score(int score, float alt_diff_frac) // hit.c:99-104
{
	if (score < 0) return score;
	score = (int)(score * (1.0 - alt_diff_frac) + .499);
	return score > 0? score : 1;
}

void hl_split_reg(mm_reg1_t *r, mm_reg1_t *r2, int n, int qlen, m128 *a, int is_qstrand) // hit.c:106-123
{
	if (n <= 0 || n >= r->cnt) return;
	*r2 = *r;
	r2->id = -1;
	r2->sam_pri = 0;
	r2->p = 0;
	r2->split_inv = 0;
	r2->cnt = r->cnt - n;
	r2->score = (int32_t)(r->score * ((float)r2->cnt / r->cnt) + .499);
	r2->as = r->as + n;
	if (r->parent == r->id) r2->parent = HL_PARENT_TMP_PRI;
	hl_reg_set_coor(r2, qlen, a, is_qstrand);
	r->cnt -= r2->cnt;
	r->score -= r2->score;
	hl_reg_set_coor(r, qlen, a, is_qstrand);
	r->split |= 1, r2->split |= 2;
}

void hl_set_parent(float mask_level, int mask_len, int n, mm_reg1_t *r, int sub_diff, int hard_mask_level, float alt_diff_frac) // hit.c:125-186
{
	if (n <= 0) return;
	for (int i = 0; i < n; ++i) r[i].id = i;
	std::vector<uint64_t> cov;
	std::vector<int> w(n);
	int k = 1;
	w[0] = 0, r[0].parent = 0;
	for (int i = 1; i < n; ++i) {
		mm_reg1_t *ri = &r[i];
		const int si = ri->qs, ei = ri->qe;
		int uncov_len = 0, j;
		bool have_uncov = hard_mask_level != 0; // hard mask level skips the uncovered-length computation
		if (!have_uncov) {
			cov.clear();
			for (j = 0; j < k; ++j) { // overlaps with existing primaries
				mm_reg1_t *rp = &r[w[j]];
				int sj = rp->qs, ej = rp->qe;
				if (ej <= si || sj >= ei) continue;
				if (sj < si) sj = si;
				if (ej > ei) ej = ei;
				cov.push_back((uint64_t)sj << 32 | (uint32_t)ej);
			}
			if (cov.empty()) { w[k++] = i, ri->parent = i, ri->n_sub = 0; continue; } // a new primary
			int x = si;
			sort64(cov);
			for (size_t c = 0; c < cov.size(); ++c) {
				if ((int)(cov[c] >> 32) > x) uncov_len += (int)(cov[c] >> 32) - x;
				x = (int32_t)cov[c] > x? (int32_t)cov[c] : x;
			}
			if (ei > x) uncov_len += ei - x;
		}
		for (j = 0; j < k; ++j) {
			mm_reg1_t *rp = &r[w[j]];
			int sj = rp->qs, ej = rp->qe, min, max, ol;
			if (ej <= si || sj >= ei) continue;
			min = ej - sj < ei - si? ej - sj : ei - si;
			max = ej - sj > ei - si? ej - sj : ei - si;
			ol = si < sj? (ei < sj? 0 : ei < ej? ei - sj : ej - sj) : (ej < si? 0 : ej < ei? ej - si : ei - si);
			if ((float)ol / min - (float)uncov_len / max > mask_level && uncov_len <= mask_len) { // secondary to rp
				int cnt_sub = 0, sci = ri->score;
				ri->parent = rp->parent;
				if (!rp->is_alt && ri->is_alt) sci = alt_score(sci, alt_diff_frac);
				rp->subsc = rp->subsc > sci? rp->subsc : sci;
				if (ri->cnt >= rp->cnt) cnt_sub = 1;
				if (rp->p && ri->p && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != min)) {
					sci = ri->p->dp_max;
					if (!rp->is_alt && ri->is_alt) sci = alt_score(sci, alt_diff_frac);
					rp->p->dp_max2 = rp->p->dp_max2 > sci? rp->p->dp_max2 : sci;
					if (rp->p->dp_max - ri->p->dp_max <= sub_diff) cnt_sub = 1;
				}
				if (cnt_sub) ++rp->n_sub;
				break;
			}
		}
		if (j == k) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
}

void hl_hit_sort(int *n_regs, mm_reg1_t *r, float alt_diff_frac) // hit.c:188-223
{
	int32_t n = *n_regs;
	if (n <= 1) return;
	std::vector<m128> aux;
	aux.reserve(n);
	for (int i = 0; i < n; ++i) {
		if (r[i].inv || r[i].cnt > 0) {
			int score = r[i].p? r[i].p->dp_max : r[i].score;
			if (r[i].is_alt) score = alt_score(score, alt_diff_frac);
			m128 e; e.x = (uint64_t)score << 32 | r[i].hash; e.y = (uint64_t)i;
			aux.push_back(e);
		} else if (r[i].p) {
			free(r[i].p);
			r[i].p = 0;
		}
	}
	sort128(aux);
	const int n_aux = (int)aux.size();
	std::vector<mm_reg1_t> t(n_aux);
	for (int i = n_aux - 1; i >= 0; --i) t[n_aux - 1 - i] = r[aux[i].y];
	if (n_aux) memcpy(r, t.data(), sizeof(mm_reg1_t) * n_aux);
	*n_regs = n_aux;
}

int hl_set_sam_pri(int n, mm_reg1_t *r) // hit.c:225-234
{
	int n_pri = 0;
	for (int i = 0; i < n; ++i)
		if (r[i].id == r[i].parent) { ++n_pri; r[i].sam_pri = (n_pri == 1); }
		else r[i].sam_pri = 0;
	return n_pri;
}

void hl_sync_regs(int n_regs, mm_reg1_t *regs) // hit.c:236-253
{
	if (n_regs <= 0) return;
	int max_id = -1;
	for (int i = 0; i < n_regs; ++i) max_id = max_id > regs[i].id? max_id : regs[i].id;
	std::vector<int> tmp(max_id + 1 > 0? max_id + 1 : 0, -1);
	for (int i = 0; i < n_regs; ++i) if (regs[i].id >= 0) tmp[regs[i].id] = i;
	for (int i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		r->id = i;
		if (r->parent == HL_PARENT_TMP_PRI) r->parent = i;
		else if (r->parent >= 0 && tmp[r->parent] >= 0) r->parent = tmp[r->parent];
		else r->parent = HL_PARENT_UNSET;
	}
	hl_set_sam_pri(n_regs, regs);
}

void hl_select_sub(float pri_ratio, int min_diff, int best_n, int check_strand, int min_strand_sc, int *n_, mm_reg1_t *r) // hit.c:255-281
{
	if (!(pri_ratio > 0.0f && *n_ > 0)) return;
	int n = *n_, n_2nd = 0, k = 0;
	std::vector<uint8_t> keep(n, 0);
	for (int i = 0; i < n; ++i) {
		int p = r[i].parent;
		if (p == i || r[i].inv) keep[i] = 1;
		else if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
			if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re))
				keep[i] = 1, ++n_2nd;
		} else if (check_strand && n_2nd < best_n && r[i].score > min_strand_sc && r[i].rev != r[p].rev) {
			r[i].strand_retained = 1;
			keep[i] = 1, ++n_2nd;
		}
	}
	for (int i = 0; i < n; ++i) {
		if (keep[i]) r[k++] = r[i];
		else if (r[i].p) free(r[i].p);
	}
	if (k != n) hl_sync_regs(k, r);
	*n_ = k;
}

int hl_filter_strand_retained(int n_regs, mm_reg1_t *r) // hit.c:283-299
{
	std::vector<uint8_t> keep(n_regs);
	int k = 0;
	for (int i = 0; i < n_regs; ++i) {
		int p = r[i].parent;
		keep[i] = (!r[i].strand_retained || r[i].div < r[p].div * 5.0f || r[i].div < 0.01f);
	}
	for (int i = 0; i < n_regs; ++i)
		if (keep[i]) { if (k < i) r[k++] = r[i]; else ++k; }
	return k;
}

void hl_filter_regs(const mm_mapopt_t *opt, int qlen, int *n_regs, mm_reg1_t *regs) // hit.c:301-320
{
	int k = 0;
	for (int i = 0; i < *n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		int flt = 0;
		if (!r->inv && !r->seg_split && r->cnt < opt->min_cnt) flt = 1;
		if (r->p) {
			if (r->mlen < opt->min_chain_score) flt = 1;
			else if (r->p->dp_max < opt->min_dp_max) flt = 1;
			else if (r->qs > qlen * opt->max_clip_ratio && qlen - r->qe > qlen * opt->max_clip_ratio) flt = 1;
			if (flt) free(r->p);
		}
		if (!flt) { if (k < i) regs[k++] = regs[i]; else ++k; }
	}
	*n_regs = k;
}

int hl_squeeze_a(int n_regs, mm_reg1_t *regs, m128 *a) // hit.c:322-341
{
	int as = 0;
	std::vector<uint64_t> aux(n_regs);
	for (int i = 0; i < n_regs; ++i) aux[i] = (uint64_t)regs[i].as << 32 | (uint32_t)i;
	sort64(aux);
	for (int i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[(int32_t)aux[i]];
		if (r->as != as) {
			memmove(&a[as], &a[r->as], (size_t)r->cnt * 16);
			r->as = as;
		}
		as += r->cnt;
	}
	return as;
}

static void set_inv_mapq(int n_regs, mm_reg1_t *regs) // hit.c:405-430
{
	if (n_regs < 3) return;
	int i;
	for (i = 0; i < n_regs; ++i) if (regs[i].inv) break;
	if (i == n_regs) return;
	std::vector<m128> aux;
	for (i = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0) {
			m128 e; e.y = (uint64_t)i; e.x = (uint64_t)regs[i].rid << 32 | (uint32_t)regs[i].rs;
			aux.push_back(e);
		}
	sort128(aux);
	const int n_aux = (int)aux.size();
	for (i = 1; i < n_aux - 1; ++i) {
		mm_reg1_t *inv = &regs[aux[i].y];
		if (inv->inv) {
			mm_reg1_t *l = &regs[aux[i-1].y], *r = &regs[aux[i+1].y];
			inv->mapq = l->mapq < r->mapq? l->mapq : r->mapq;
		}
	}
}

void hl_set_mapq(int n_regs, mm_reg1_t *regs, int min_chain_sc, int match_sc, int rep_len, int is_sr, int is_splice) // hit.c:432-485
{
	static const float q_coef = 40.0f;
	int64_t sum_sc = 0;
	int n_2nd_splice = 0;
	if (n_regs == 0) return;
	for (int i = 0; i < n_regs; ++i) {
		if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
		else if (regs[i].is_spliced) ++n_2nd_splice;
	}
	const float uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
	for (int i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		if (r->inv) r->mapq = 0;
		else if (r->parent == r->id) {
			int mapq, subsc;
			float pen_s1 = (r->score > 100? 1.0f : 0.01f * r->score) * uniq_ratio;
			float pen_cm = r->cnt > 10? 1.0f : 0.1f * r->cnt;
			pen_cm = pen_s1 < pen_cm? pen_s1 : pen_cm;
			subsc = r->subsc > min_chain_sc? r->subsc : min_chain_sc;
			if (r->p && r->p->dp_max2 > 0 && r->p->dp_max > 0) {
				float x, identity = (float)r->mlen / r->blen;
				if (is_sr && is_splice) x = (float)r->p->dp_max2 / r->p->dp_max;
				else x = (float)r->p->dp_max2 * subsc / r->p->dp_max / r->score0;
				mapq = (int)(identity * pen_cm * q_coef * (1.0f - x * x) * logf((float)r->p->dp_max / match_sc));
				if (!is_sr) {
					int mapq_alt = (int)(6.02f * identity * identity * (r->p->dp_max - r->p->dp_max2) / match_sc + .499f);
					mapq = mapq < mapq_alt? mapq : mapq_alt;
				}
				if (is_splice && is_sr && r->is_spliced && n_2nd_splice == 0) mapq += 10;
			} else {
				float x = (float)subsc / r->score0;
				if (r->p) {
					float identity = (float)r->mlen / r->blen;
					mapq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf((float)r->p->dp_max / match_sc));
				} else mapq = (int)(pen_cm * q_coef * (1.0f - x) * logf(r->score));
			}
			mapq -= (int)(4.343f * logf(r->n_sub + 1) + .499f);
			mapq = mapq > 0? mapq : 0;
			r->mapq = mapq < 60? mapq : 60;
			if (r->p && r->p->dp_max > r->p->dp_max2 && r->mapq == 0) r->mapq = 1;
		} else r->mapq = 0;
	}
	set_inv_mapq(n_regs, regs);
}

// ---- esterr.c:5-64 ----
static inline int32_t for_qpos(int32_t qlen, const m128 *a)
{
	int32_t x = (int32_t)a->y, q_span = (int32_t)(a->y >> 32 & 0xff);
	if (a->x >> 63) x = qlen - 1 - (x + 1 - q_span);
	return x;
}

static int mini_idx(int qlen, const m128 *a, int32_t n, const uint64_t *mini_pos)
{
	int32_t x = for_qpos(qlen, a), L = 0, R = n - 1;
	while (L <= R) {
		int32_t m = (int32_t)(((uint64_t)L + R) >> 1), y = (int32_t)mini_pos[m];
		if (y < x) L = m + 1;
		else if (y > x) R = m - 1;
		else return m;
	}
	return -1;
}

void hl_est_err(const mm_idx_t *mi, int qlen, int n_regs, mm_reg1_t *regs, const m128 *a, int32_t n, const uint64_t *mini_pos)
{
	if (n == 0) return;
	uint64_t sum_k = 0;
	for (int i = 0; i < n; ++i) sum_k += mini_pos[i] >> 32 & 0xff;
	const float avg_k = (float)sum_k / n;
	for (int i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		int32_t st, en, j, k, n_match, n_tot, l_ref;
		r->div = -1.0f;
		if (r->cnt == 0) continue;
		st = en = mini_idx(qlen, r->rev? &a[r->as + r->cnt - 1] : &a[r->as], n, mini_pos);
		if (st < 0) {
			if (mm_verbose >= 2) fprintf(stderr, "[WARNING] logic inconsistency in mm_est_err(). Please contact the developer.\n");
			continue;
		}
		l_ref = mi->seq[r->rid].len;
		for (k = 1, j = st + 1, n_match = 1; j < n && k < r->cnt; ++j) {
			int32_t x = for_qpos(qlen, r->rev? &a[r->as + r->cnt - 1 - k] : &a[r->as + k]);
			if (x == (int32_t)mini_pos[j]) ++k, en = j, ++n_match;
		}
		n_tot = en - st + 1;
		if (r->qs > avg_k && r->rs > avg_k) ++n_tot;
		if (qlen - r->qs > avg_k && l_ref - r->re > avg_k) ++n_tot;
		r->div = n_match >= n_tot? 0.0f : (float)(1.0 - pow((double)n_match / n_tot, 1.0 / avg_k));
	}
}

// ---- align.c:977-1046 (event identity, dp_max re-ranking) ----
static inline void count_gaps(const mm_reg1_t *r, int32_t *n_gap_, int32_t *n_gapo_)
{
	int32_t n_gapo = 0, n_gap = 0;
	*n_gap_ = *n_gapo_ = -1;
	if (r->p == 0) return;
	for (uint32_t i = 0; i < r->p->n_cigar; ++i) {
		int32_t op = r->p->cigar[i] & 0xf, len = r->p->cigar[i] >> 4;
		if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) ++n_gapo, n_gap += len;
	}
	*n_gap_ = n_gap, *n_gapo_ = n_gapo;
}

double hl_event_identity(const mm_reg1_t *r)
{
	int32_t n_gap, n_gapo;
	if (r->p == 0) return -1.0f;
	count_gaps(r, &n_gap, &n_gapo);
	return (double)r->mlen / (r->blen + r->p->n_ambi - n_gap + n_gapo);
}

static int32_t recal_max_dp(const mm_reg1_t *r, double b2, int32_t match_sc)
{
	int32_t n_gap = 0, n_mis;
	double gap_cost = 0.0;
	if (r->p == 0) return -1;
	for (uint32_t i = 0; i < r->p->n_cigar; ++i) {
		int32_t op = r->p->cigar[i] & 0xf, len = r->p->cigar[i] >> 4;
		if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
			gap_cost += b2 + (double)mmx_log2(1.0 + len);
			n_gap += len;
		}
	}
	n_mis = r->blen + r->p->n_ambi - r->mlen - n_gap;
	return (int32_t)(match_sc * (r->mlen - b2 * n_mis - gap_cost) + .499);
}

void hl_update_dp_max(int qlen, int n_regs, mm_reg1_t *regs, float frac, int a, int b)
{
	int32_t max = -1, max2 = -1, max_i = -1;
	if (n_regs < 2) return;
	for (int i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		if (r->p == 0) continue;
		if (r->p->dp_max > max) max2 = max, max = r->p->dp_max, max_i = i;
		else if (r->p->dp_max > max2) max2 = r->p->dp_max;
	}
	if (max_i < 0 || max < 0 || max2 < 0) return;
	if (regs[max_i].qe - regs[max_i].qs < (double)qlen * frac) return;
	if (max2 < (double)max * frac) return;
	double div = 1. - hl_event_identity(&regs[max_i]);
	if (div < 0.02) div = 0.02;
	double b2 = 0.5 / div;
	if (b2 * a < b) b2 = (double)a / b;
	for (int i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[i];
		if (r->p == 0) continue;
		r->p->dp_max = recal_max_dp(r, b2, a);
		if (r->p->dp_max < 0) r->p->dp_max = 0;
	}
}


// ---- symmetric DUST (sdust.c:66-186, W-window low-complexity masking): used by -T to drop query minimizers that lie mostly in
// masked intervals (mm_dust_minier, map.c:33-57). Restated on std containers; the state handling is the reference's, including that
// an ambiguous base ends the run (l, t) but leaves the word window and its counters as they are. ----
#include <deque>
namespace {
struct DustState {
	static constexpr int WLEN = 3, WTOT = 1 << (WLEN << 1);
	struct Perf { int start, finish, r, l; };
	std::deque<int> w;          // triplet words of the current window
	std::vector<Perf> P;        // perfect intervals of the window, by descending start then ascending finish
	std::vector<uint64_t> &res; // masked intervals start<<32 | finish
	int cv[WTOT], cw[WTOT], rv = 0, rw = 0, L = 0;
	const int T, W;
	DustState(std::vector<uint64_t> &out, int T_, int W_) : res(out), T(T_), W(W_) { memset(cv, 0, sizeof(cv)); memset(cw, 0, sizeof(cw)); }

	void shift_window(int t) { // sdust.c:66-88
		int s;
		if ((int)w.size() >= W - WLEN + 1) {
			s = w.front(); w.pop_front();
			rw -= --cw[s];
			if (L > (int)w.size()) --L, rv -= --cv[s];
		}
		w.push_back(t);
		++L;
		rw += cw[t]++;
		rv += cv[t]++;
		if (cv[t] * 10 > T << 1) {
			do {
				s = w[w.size() - L];
				rv -= --cv[s];
				--L;
			} while (s != t);
		}
	}
	void save_masked(int start) { // sdust.c:90-104
		if (P.empty() || P.back().start >= start) return;
		const Perf &p = P.back();
		bool saved = false;
		if (!res.empty()) {
			const int s = (int)(res.back() >> 32), f = (int)(uint32_t)res.back();
			if (p.start <= f) saved = true, res.back() = (uint64_t)s << 32 | (uint32_t)(f > p.finish? f : p.finish);
		}
		if (!saved) res.push_back((uint64_t)p.start << 32 | (uint32_t)p.finish);
		int i = (int)P.size() - 1;
		while (i >= 0 && P[i].start < start) --i;
		P.resize(i + 1);
	}
	void find_perfect(int start) { // sdust.c:106-133
		int c[WTOT], r = rv, max_r = 0, max_l = 0;
		memcpy(c, cv, sizeof(c));
		for (int i = (int)w.size() - L - 1; i >= 0; --i) {
			const int t = w[i];
			r += c[t]++;
			const int new_r = r, new_l = (int)w.size() - i - 1;
			if (new_r * 10 > T * new_l) {
				size_t j;
				for (j = 0; j < P.size() && P[j].start >= i + start; ++j)
					if (max_r == 0 || P[j].r * max_l > max_r * P[j].l) max_r = P[j].r, max_l = P[j].l;
				if (max_r == 0 || new_r * max_l >= max_r * new_l) {
					max_r = new_r, max_l = new_l;
					P.insert(P.begin() + j, Perf{ i + start, (int)w.size() + (WLEN - 1) + start, new_r, new_l });
				}
			}
		}
	}
};
} // namespace

void hl_sdust(const uint8_t *seq, int l_seq, int T, int W, std::vector<uint64_t> &res) // sdust_core (sdust.c:135-170); seq is ASCII
{
	res.clear();
	DustState d(res, T, W);
	int l = 0, start;
	unsigned t = 0;
	for (int i = 0; i <= l_seq; ++i) {
		const int b = i < l_seq? mmx_nt4(seq[i]) : 4;
		if (b < 4) {
			++l, t = (t << 2 | (unsigned)b) & (DustState::WTOT - 1);
			if (l >= DustState::WLEN) {
				start = (l - W > 0? l - W : 0) + (i + 1 - l);
				d.save_masked(start);
				d.shift_window((int)t);
				if (d.rw * 10 > d.L * T) d.find_perfect(start);
			}
		} else {
			start = (l - W + 1 > 0? l - W + 1 : 0) + (i + 1 - l);
			while (!d.P.empty()) d.save_masked(start++);
			l = 0, t = 0;
		}
	}
}
