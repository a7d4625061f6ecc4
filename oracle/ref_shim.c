/* oracle/ref_shim.c -- TEST INFRASTRUCTURE. Thin exports over the UNMODIFIED reference so that tests can
 * reach per-stage results without patching it. Compiled only into oracle/_ref/libminimap2_ref.so. */
#include <stdlib.h>
#include <string.h>
#include "minimap.h"
#include "mmpriv.h"
#include "kalloc.h"
#include "ksw2.h"

/* size/offset probes so the ctypes mirrors in tests/ can assert ABI identity with include/minimap.h */
int refshim_sizeof_mapopt(void) { return (int)sizeof(mm_mapopt_t); }
int refshim_sizeof_idxopt(void) { return (int)sizeof(mm_idxopt_t); }
int refshim_sizeof_reg1(void)   { return (int)sizeof(mm_reg1_t); }
int refshim_sizeof_idx(void)    { return (int)sizeof(mm_idx_t); }
int refshim_sizeof_extra(void)  { return (int)sizeof(mm_extra_t); }

/* mm_sketch with a malloc'd result the caller frees with refshim_free() */
int refshim_sketch(const char *seq, int len, int w, int k, uint32_t rid, int is_hpc, mm128_t **out)
{
	mm128_v v = {0,0,0};
	mm_sketch(0, seq, len, w, k, rid, is_hpc, &v);
	*out = v.a;
	return (int)v.n;
}
void refshim_free(void *p) { free(p); }

/* mg_lchain_dp / mg_lchain_rmq on a caller array (copied, because the reference frees its input) */
int refshim_lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
					  float pen_gap, float pen_skip, int is_cdna, int n_seg, int64_t n, const mm128_t *a_in,
					  uint64_t **u_out, mm128_t **a_out, int *n_a_out)
{
	void *km = km_init();
	mm128_t *a = (mm128_t*)kmalloc(km, (n > 0? n : 1) * sizeof(mm128_t)), *b;
	uint64_t *u = 0;
	int n_u = 0, i, tot = 0;
	memcpy(a, a_in, n * sizeof(mm128_t));
	b = mg_lchain_dp(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, pen_gap, pen_skip, is_cdna, n_seg, n, a, &n_u, &u, km);
	for (i = 0; i < n_u; ++i) tot += (int32_t)u[i];
	*u_out = (uint64_t*)malloc((n_u > 0? n_u : 1) * 8);
	*a_out = (mm128_t*)malloc((tot > 0? tot : 1) * 16);
	if (n_u > 0) memcpy(*u_out, u, n_u * 8), memcpy(*a_out, b, tot * 16);
	*n_a_out = tot;
	kfree(km, b); kfree(km, u);
	km_destroy(km);
	return n_u;
}

int refshim_lchain_rmq(int max_dist, int max_dist_inner, int bw, int max_skip, int cap_rmq_size, int min_cnt, int min_sc,
					   float pen_gap, float pen_skip, int64_t n, const mm128_t *a_in,
					   uint64_t **u_out, mm128_t **a_out, int *n_a_out)
{
	void *km = km_init();
	mm128_t *a = (mm128_t*)kmalloc(km, (n > 0? n : 1) * sizeof(mm128_t)), *b;
	uint64_t *u = 0;
	int n_u = 0, i, tot = 0;
	memcpy(a, a_in, n * sizeof(mm128_t));
	b = mg_lchain_rmq(max_dist, max_dist_inner, bw, max_skip, cap_rmq_size, min_cnt, min_sc, pen_gap, pen_skip, n, a, &n_u, &u, km);
	for (i = 0; i < n_u; ++i) tot += (int32_t)u[i];
	*u_out = (uint64_t*)malloc((n_u > 0? n_u : 1) * 8);
	*a_out = (mm128_t*)malloc((tot > 0? tot : 1) * 16);
	if (n_u > 0) memcpy(*u_out, u, n_u * 8), memcpy(*a_out, b, tot * 16);
	*n_a_out = tot;
	kfree(km, b); kfree(km, u);
	km_destroy(km);
	return n_u;
}

/* ksw_extd2_sse with libc allocation (km=0); the caller frees ez->cigar with refshim_free() */
void refshim_extd2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int8_t m, const int8_t *mat,
				   int8_t gapo, int8_t gape, int8_t gapo2, int8_t gape2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	memset(ez, 0, sizeof(*ez));
	ksw_extd2_sse(0, qlen, q, tlen, t, m, mat, gapo, gape, gapo2, gape2, w, zdrop, end_bonus, flag, ez);
}
void refshim_extz2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int8_t m, const int8_t *mat,
				   int8_t gapo, int8_t gape, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez)
{
	memset(ez, 0, sizeof(*ez));
	ksw_extz2_sse(0, qlen, q, tlen, t, m, mat, gapo, gape, w, zdrop, end_bonus, flag, ez);
}
void refshim_exts2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int8_t m, const int8_t *mat, int8_t gapo, int8_t gape, int8_t gapo2,
				   int8_t noncan, int zdrop, int end_bonus, int8_t junc_bonus, int8_t junc_pen, int flag, const uint8_t *junc, ksw_extz_t *ez)
{
	memset(ez, 0, sizeof(*ez));
	ksw_exts2_sse(0, qlen, q, tlen, t, m, mat, gapo, gape, gapo2, noncan, zdrop, end_bonus, junc_bonus, junc_pen, flag, junc, ez);
}
int refshim_ll_i16(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int m, const int8_t *mat, int gapo, int gape, int *qe, int *te)
{
	void *qp = ksw_ll_qinit(0, 2, qlen, q, m, mat);
	int sc = ksw_ll_i16(qp, tlen, t, gapo, gape, qe, te);
	free(qp);
	return sc;
}

/* index probe: occurrences of one minimizer hash */
int refshim_idx_get(const mm_idx_t *mi, uint64_t minier, uint64_t *out, int max_out)
{
	int n, i;
	const uint64_t *cr = mm_idx_get(mi, minier, &n);
	for (i = 0; i < n && i < max_out; ++i) out[i] = cr[i];
	return n;
}
void refshim_sort128x(mm128_t *a, int64_t n) { radix_sort_128x(a, a + n); }
void refshim_sort64(uint64_t *a, int64_t n) { radix_sort_64(a, a + n); }
