/* oracle/mm2o_chain.c -- TEST INFRASTRUCTURE ONLY (see mm2o.h).
 * Restates mg_lchain_dp (lchain.c:148-217), comput_sc (lchain.c:113-138), mg_log2 (mmpriv.h:139-147),
 * mg_chain_backtrack / mg_chain_bk_end (lchain.c:9-76) and compact_a (lchain.c:78-111).
 * Compiled with -ffp-contract=off: the float penalty must round like the reference's x86-64 -O2 build (no FMA). */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "mm2o.h"

#define SEG_SHIFT 48
#define SEG_MASK  (0xffULL << SEG_SHIFT)

float mm2o_log2(float x) /* mmpriv.h:139-147; valid for x >= 2 */
{
	union { float f; uint32_t i; } z = { x };
	float r = (float)((z.i >> 23) & 255) - 128;
	z.i &= ~(255U << 23);
	z.i += 127U << 23;
	r += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return r;
}

static int32_t pair_score(const mm2o_128_t *ai, const mm2o_128_t *aj, int32_t max_dist_x, int32_t max_dist_y, int32_t bw,
						  float pen_gap, float pen_skip, int is_cdna, int n_seg) /* lchain.c:113-138 */
{
	int32_t dq = (int32_t)ai->y - (int32_t)aj->y, dr, dd, dg, span_j, sc;
	int32_t si = (int32_t)((ai->y & SEG_MASK) >> SEG_SHIFT), sj = (int32_t)((aj->y & SEG_MASK) >> SEG_SHIFT);
	if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
	dr = (int32_t)(ai->x - aj->x);
	if (si == sj && (dr == 0 || dq > max_dist_y)) return INT32_MIN;
	dd = dr > dq? dr - dq : dq - dr;
	if (si == sj && dd > bw) return INT32_MIN;
	if (n_seg > 1 && !is_cdna && si == sj && dr > max_dist_y) return INT32_MIN;
	dg = dr < dq? dr : dq;
	span_j = (int32_t)(aj->y >> 32 & 0xff);
	sc = span_j < dg? span_j : dg;
	if (dd || dg > span_j) {
		float lin = pen_gap * (float)dd + pen_skip * (float)dg;
		float lg = dd >= 1? mm2o_log2((float)(dd + 1)) : 0.0f;
		if (is_cdna || si != sj) {
			if (si != sj && dr == 0) ++sc;
			else if (dr > dq || si != sj) sc -= (int)(lin < lg? lin : lg);
			else sc -= (int)(lin + .5f * lg);
		} else sc -= (int)(lin + .5f * lg);
	}
	return sc;
}

static int64_t walk_end(int32_t max_drop, const mm2o_128_t *z, const int32_t *f, const int64_t *p, int32_t *t, int64_t k) /* lchain.c:9-25 */
{
	int64_t i = (int64_t)z[k].y, end_i = -1, max_i = i;
	int32_t max_s = 0;
	if (i < 0 || t[i] != 0) return i;
	do {
		int32_t s;
		t[i] = 2;
		end_i = i = p[i];
		s = i < 0? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
		if (s > max_s) max_s = s, max_i = i;
		else if (max_s - s > max_drop) break;
	} while (i >= 0 && t[i] == 0);
	for (i = (int64_t)z[k].y; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
	return max_i;
}

/* lchain.c:27-76; fills v[] (anchor indices, chain by chain, last anchor first) and returns u[] */
uint64_t *mm2o_chain_backtrack(int64_t n, const int32_t *f, const int64_t *p, int32_t *v, int32_t *t,
							   int32_t min_cnt, int32_t min_sc, int32_t max_drop, int32_t *n_u_, int32_t *n_v_)
{
	mm2o_128_t *z;
	uint64_t *u = 0;
	int64_t i, k, n_z = 0, n_v;
	int32_t n_u, pass;
	*n_u_ = *n_v_ = 0;
	for (i = 0; i < n; ++i) if (f[i] >= min_sc) ++n_z;
	if (n_z == 0) return 0;
	z = (mm2o_128_t*)malloc(n_z * sizeof(mm2o_128_t));
	for (i = 0, k = 0; i < n; ++i) if (f[i] >= min_sc) z[k].x = f[i], z[k++].y = i;
	mm2o_radix_sort_128x(z, z + n_z);
	for (pass = 0; pass < 2; ++pass) { /* pass 0 counts n_u, pass 1 fills (identical traversals in the reference) */
		memset(t, 0, n * 4);
		for (k = n_z - 1, n_v = n_u = 0; k >= 0; --k) {
			if (t[z[k].y] == 0) {
				int64_t n_v0 = n_v, end_i = walk_end(max_drop, z, f, p, t, k);
				int32_t sc;
				for (i = (int64_t)z[k].y; i != end_i; i = p[i]) {
					if (pass) v[n_v] = (int32_t)i;
					++n_v, t[i] = 1;
				}
				sc = i < 0? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
				if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) {
					if (pass) u[n_u] = (uint64_t)sc << 32 | (uint64_t)(n_v - n_v0);
					++n_u;
				} else n_v = n_v0;
			}
		}
		if (pass == 0) u = (uint64_t*)malloc((n_u > 0? n_u : 1) * 8);
	}
	free(z);
	*n_u_ = n_u, *n_v_ = (int32_t)n_v;
	return u;
}

/* lchain.c:78-111; returns the compacted anchor array (malloc'd), reorders u[] in place */
mm2o_128_t *mm2o_compact(int32_t n_u, uint64_t *u, int32_t n_v, const int32_t *v, const mm2o_128_t *a)
{
	mm2o_128_t *b, *w, *out;
	uint64_t *u2;
	int64_t i, j, k;
	b = (mm2o_128_t*)malloc((n_v > 0? n_v : 1) * sizeof(mm2o_128_t));
	for (i = 0, k = 0; i < n_u; ++i) {
		int32_t k0 = (int32_t)k, ni = (int32_t)u[i];
		for (j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
	}
	w = (mm2o_128_t*)malloc(n_u * sizeof(mm2o_128_t));
	for (i = k = 0; i < n_u; ++i) {
		w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | (uint64_t)i;
		k += (int32_t)u[i];
	}
	mm2o_radix_sort_128x(w, w + n_u);
	u2 = (uint64_t*)malloc(n_u * 8);
	out = (mm2o_128_t*)malloc((n_v > 0? n_v : 1) * sizeof(mm2o_128_t));
	for (i = k = 0; i < n_u; ++i) {
		int32_t j2 = (int32_t)w[i].y, nn = (int32_t)u[j2];
		u2[i] = u[j2];
		memcpy(&out[k], &b[w[i].y >> 32], nn * sizeof(mm2o_128_t));
		k += nn;
	}
	memcpy(u, u2, n_u * 8);
	free(b); free(w); free(u2);
	return out;
}

int mm2o_lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
				   float pen_gap, float pen_skip, int is_cdna, int n_seg, int64_t n, const mm2o_128_t *a,
				   uint64_t **u_out, mm2o_128_t **a_out, int *n_a_out)
{
	int32_t *f, *t, *v, n_u, n_v, max_drop = bw;
	int64_t *p, i, j, best_prev = -1, st = 0;
	uint64_t *u;
	*u_out = 0, *a_out = 0, *n_a_out = 0;
	if (n == 0 || a == 0) return 0;
	if (max_dist_x < bw) max_dist_x = bw;
	if (max_dist_y < bw && !is_cdna) max_dist_y = bw;
	if (is_cdna) max_drop = INT32_MAX;
	p = (int64_t*)malloc(n * 8);
	f = (int32_t*)malloc(n * 4), v = (int32_t*)malloc(n * 4), t = (int32_t*)calloc(n, 4);
	for (i = 0; i < n; ++i) { /* lchain.c:168-207 */
		int64_t max_j = -1, end_j;
		int32_t max_f = (int32_t)(a[i].y >> 32 & 0xff), n_skip = 0;
		while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + max_dist_x)) ++st;
		if (i - st > max_iter) st = i - max_iter;
		for (j = i - 1; j >= st; --j) {
			int32_t sc = pair_score(&a[i], &a[j], max_dist_x, max_dist_y, bw, pen_gap, pen_skip, is_cdna, n_seg);
			if (sc == INT32_MIN) continue;
			sc += f[j];
			if (sc > max_f) {
				max_f = sc, max_j = j;
				if (n_skip > 0) --n_skip;
			} else if (t[j] == (int32_t)i) {
				if (++n_skip > max_skip) break;
			}
			if (p[j] >= 0) t[p[j]] = (int32_t)i;
		}
		end_j = j;
		if (best_prev < 0 || a[i].x - a[best_prev].x > (uint64_t)(int64_t)max_dist_x) { /* NB: reference compares uint64 - uint64 > (int64_t)max_dist_x, i.e. unsigned */
			int32_t mx = INT32_MIN;
			best_prev = -1;
			for (j = i - 1; j >= st; --j) if (mx < f[j]) mx = f[j], best_prev = j;
		}
		if (best_prev >= 0 && best_prev < end_j) {
			int32_t tmp = pair_score(&a[i], &a[best_prev], max_dist_x, max_dist_y, bw, pen_gap, pen_skip, is_cdna, n_seg);
			if (tmp != INT32_MIN && max_f < tmp + f[best_prev]) max_f = tmp + f[best_prev], max_j = best_prev;
		}
		f[i] = max_f, p[i] = max_j;
		v[i] = max_j >= 0 && v[max_j] > max_f? v[max_j] : max_f;
		if (best_prev < 0 || (a[i].x - a[best_prev].x <= (uint64_t)(int64_t)max_dist_x && f[best_prev] < f[i])) best_prev = i;
	}
	u = mm2o_chain_backtrack(n, f, p, v, t, min_cnt, min_sc, max_drop, &n_u, &n_v);
	free(p); free(f); free(t);
	if (n_u == 0) { free(v); free(u); return 0; }
	*a_out = mm2o_compact(n_u, u, n_v, v, a);
	*u_out = u, *n_a_out = n_v;
	free(v);
	return n_u;
}
