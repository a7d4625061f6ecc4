/* oracle/mm2o_rmq.c -- TEST INFRASTRUCTURE ONLY (see mm2o.h).
 * Restates mg_lchain_rmq (lchain.c:250-368) with comput_sc_simple (lchain.c:232-248) and the instantiation of krmq.h
 * used there (lchain.c:219-230: key = (y, i), "less" on the priority `pri`).
 *
 * The range-minimum query of krmq.h returns the first minimum it meets on the two root paths, so on equal priorities the
 * answer depends on the SHAPE of the AVL tree. The tree below therefore follows krmq.h's insertion (:197-244), deletion
 * (:246-330), rotations (:160-195), subtree-minimum maintenance (:155-158), query (:110-151), interval (:96-108) and
 * backward iterator (:345-385) step by step. Nodes are array slots: slot j is anchor j (an anchor is in a tree at most
 * once), slot n is the stand-in for krmq.h's stack-allocated `fake` root; links are indices, -1 is NULL.
 * Compiled with -ffp-contract=off like mm2o_chain.c. */
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "mm2o.h"

float mm2o_log2(float x);
uint64_t *mm2o_chain_backtrack(int64_t n, const int32_t *f, const int64_t *p, int32_t *v, int32_t *t,
							   int32_t min_cnt, int32_t min_sc, int32_t max_drop, int32_t *n_u_, int32_t *n_v_);
mm2o_128_t *mm2o_compact(int32_t n_u, uint64_t *u, int32_t n_v, const int32_t *v, const mm2o_128_t *a);

#define NIL   (-1)
#define DEPTH 64

typedef struct {
	int32_t *kid[2];  /* children */
	int32_t *best;    /* slot with the smallest priority in the subtree (krmq.h: `s`) */
	signed char *bal;
	uint32_t *size;
	int32_t *y;
	double *pri;
	int32_t root, fake;
} tree_t;

static void tree_init(tree_t *T, int64_t n)
{
	T->kid[0] = (int32_t*)malloc((n + 1) * 4), T->kid[1] = (int32_t*)malloc((n + 1) * 4);
	T->best = (int32_t*)malloc((n + 1) * 4), T->bal = (signed char*)malloc(n + 1), T->size = (uint32_t*)malloc((n + 1) * 4);
	T->y = (int32_t*)malloc((n + 1) * 4), T->pri = (double*)malloc((n + 1) * 8);
	T->root = NIL, T->fake = (int32_t)n;
}
static void tree_free(tree_t *T)
{
	free(T->kid[0]); free(T->kid[1]); free(T->best); free(T->bal); free(T->size); free(T->y); free(T->pri);
}

/* lc_elem_cmp (lchain.c:225): by y, then by i */
static int key_cmp(int32_t ay, int64_t ai, const tree_t *T, int32_t b)
{
	if (ay != T->y[b]) return ay < T->y[b]? -1 : 1;
	return (ai > b) - (ai < b);
}
#define PRI_LT(T, a, b) ((T)->pri[a] < (T)->pri[b])
static uint32_t kid_size(const tree_t *T, int32_t p, int d) { return T->kid[d][p] == NIL? 0 : T->size[T->kid[d][p]]; }

static void fix_best(tree_t *T, int32_t p, int32_t q, int32_t r) /* krmq.h:155-158 */
{
	T->best[p] = (q == NIL || PRI_LT(T, p, T->best[q]))? p : T->best[q];
	T->best[p] = (r == NIL || PRI_LT(T, T->best[p], T->best[r]))? T->best[p] : T->best[r];
}

static int32_t rot1(tree_t *T, int32_t p, int dir) /* krmq.h:160-172 */
{
	int opp = 1 - dir;
	int32_t q = T->kid[opp][p], s = T->best[p];
	uint32_t sz = T->size[p];
	T->size[p] -= T->size[q] - kid_size(T, q, dir);
	T->size[q] = sz;
	fix_best(T, p, T->kid[dir][p], T->kid[dir][q]);
	T->best[q] = s;
	T->kid[opp][p] = T->kid[dir][q];
	T->kid[dir][q] = p;
	return q;
}

static int32_t rot2(tree_t *T, int32_t p, int dir) /* krmq.h:174-195 */
{
	int opp = 1 - dir, b1;
	int32_t q = T->kid[opp][p], r = T->kid[dir][q], s = T->best[p];
	uint32_t sz_x = kid_size(T, r, dir);
	T->size[r] = T->size[p];
	T->size[p] -= T->size[q] - sz_x;
	T->size[q] -= sz_x + 1;
	fix_best(T, p, T->kid[dir][p], T->kid[dir][r]);
	fix_best(T, q, T->kid[opp][q], T->kid[opp][r]);
	T->best[r] = s;
	T->kid[opp][p] = T->kid[dir][r];
	T->kid[dir][r] = p;
	T->kid[dir][q] = T->kid[opp][r];
	T->kid[opp][r] = q;
	b1 = dir == 0? 1 : -1;
	if (T->bal[r] == b1) T->bal[q] = 0, T->bal[p] = (signed char)-b1;
	else if (T->bal[r] == 0) T->bal[q] = T->bal[p] = 0;
	else T->bal[q] = (signed char)b1, T->bal[p] = 0;
	T->bal[r] = 0;
	return r;
}

static void tree_insert(tree_t *T, int32_t x) /* krmq.h:197-244; y[x], pri[x] already set */
{
	unsigned char turn[DEPTH];
	int32_t way[DEPTH], bp = T->root, bq = NIL, p, q, r;
	int i, which = 0, top = 0, n_way = 0, b1;
	for (p = bp, q = bq; p != NIL; q = p, p = T->kid[which][p]) {
		int c = key_cmp(T->y[x], x, T, p);
		if (c == 0) return;
		if (T->bal[p] != 0) bq = q, bp = p, top = 0;
		turn[top++] = (unsigned char)(which = c > 0);
		way[n_way++] = p;
	}
	T->bal[x] = 0, T->size[x] = 1, T->kid[0][x] = T->kid[1][x] = NIL, T->best[x] = x;
	if (q == NIL) T->root = x;
	else T->kid[which][q] = x;
	if (bp == NIL) return;
	for (i = 0; i < n_way; ++i) ++T->size[way[i]];
	for (i = n_way - 1; i >= 0; --i) {
		fix_best(T, way[i], T->kid[0][way[i]], T->kid[1][way[i]]);
		if (T->best[way[i]] != x) break;
	}
	for (p = bp, top = 0; p != x; p = T->kid[turn[top]][p], ++top) {
		if (turn[top] == 0) --T->bal[p];
		else ++T->bal[p];
	}
	if (T->bal[bp] > -2 && T->bal[bp] < 2) return;
	which = T->bal[bp] < 0;
	b1 = which == 0? 1 : -1;
	q = T->kid[1 - which][bp];
	if (T->bal[q] == b1) {
		r = rot1(T, bp, which);
		T->bal[q] = T->bal[bp] = 0;
	} else r = rot2(T, bp, which);
	if (bq == NIL) T->root = r;
	else T->kid[bp != T->kid[0][bq]][bq] = r;
}

static int32_t tree_find(const tree_t *T, int32_t ky, int64_t ki) /* krmq.h:83-95 */
{
	int32_t p = T->root;
	while (p != NIL) {
		int c = key_cmp(ky, ki, T, p);
		if (c == 0) break;
		p = T->kid[c > 0][p];
	}
	return p;
}

static void tree_erase(tree_t *T, int32_t ky, int64_t ki) /* krmq.h:246-330 */
{
	int32_t p, way[DEPTH], fake = T->fake;
	unsigned char turn[DEPTH];
	int i, d = 0, c;
	if (T->root == NIL) return;
	T->y[fake] = T->y[T->root], T->pri[fake] = T->pri[T->root], T->best[fake] = T->best[T->root];
	T->bal[fake] = T->bal[T->root], T->size[fake] = T->size[T->root];
	T->kid[0][fake] = T->root, T->kid[1][fake] = NIL;
	for (c = -1, p = fake; c; c = key_cmp(ky, ki, T, p)) {
		int which = c > 0;
		turn[d] = (unsigned char)which;
		way[d++] = p;
		p = T->kid[which][p];
		if (p == NIL) return;
	}
	for (i = 1; i < d; ++i) --T->size[way[i]];
	if (T->kid[1][p] == NIL) {
		T->kid[turn[d-1]][way[d-1]] = T->kid[0][p];
	} else {
		int32_t q = T->kid[1][p];
		if (T->kid[0][q] == NIL) {
			T->kid[0][q] = T->kid[0][p];
			T->bal[q] = T->bal[p];
			T->kid[turn[d-1]][way[d-1]] = q;
			way[d] = q, turn[d++] = 1;
			T->size[q] = T->size[p] - 1;
		} else {
			int32_t r;
			int e = d++;
			for (;;) {
				turn[d] = 0;
				way[d++] = q;
				r = T->kid[0][q];
				if (T->kid[0][r] == NIL) break;
				q = r;
			}
			T->kid[0][r] = T->kid[0][p];
			T->kid[0][q] = T->kid[1][r];
			T->kid[1][r] = T->kid[1][p];
			T->bal[r] = T->bal[p];
			T->kid[turn[e-1]][way[e-1]] = r;
			way[e] = r, turn[e] = 1;
			for (i = e + 1; i < d; ++i) --T->size[way[i]];
			T->size[r] = T->size[p] - 1;
		}
	}
	for (i = d - 1; i >= 0; --i) fix_best(T, way[i], T->kid[0][way[i]], T->kid[1][way[i]]);
	while (--d > 0) {
		int32_t q = way[d];
		int which = turn[d], other = 1 - which, b1 = 1, b2 = 2;
		if (which) b1 = -b1, b2 = -b2;
		T->bal[q] += (signed char)b1;
		if (T->bal[q] == b1) break;
		else if (T->bal[q] == b2) {
			int32_t r = T->kid[other][q];
			if (T->bal[r] == -b1) {
				T->kid[turn[d-1]][way[d-1]] = rot2(T, q, which);
			} else {
				T->kid[turn[d-1]][way[d-1]] = rot1(T, q, which);
				if (T->bal[r] == 0) {
					T->bal[r] = (signed char)-b1;
					T->bal[q] = (signed char)b1;
					break;
				} else T->bal[r] = T->bal[q] = 0;
			}
		}
	}
	T->root = T->kid[0][fake];
}

/* krmq.h:110-151: minimum priority over keys in [lo, up] */
static int32_t tree_rmq(const tree_t *T, int32_t lo_y, int64_t lo_i, int32_t up_y, int64_t up_i)
{
	int32_t p, way[2][DEPTH], mn;
	signed char c[2][DEPTH];
	int len[2] = {0, 0}, i, lca;
	if (T->root == NIL) return NIL;
	for (p = T->root; p != NIL;) {
		int cc = key_cmp(lo_y, lo_i, T, p);
		way[0][len[0]] = p, c[0][len[0]++] = (signed char)cc;
		if (cc == 0) break;
		p = T->kid[cc > 0][p];
	}
	for (p = T->root; p != NIL;) {
		int cc = key_cmp(up_y, up_i, T, p);
		way[1][len[1]] = p, c[1][len[1]++] = (signed char)cc;
		if (cc == 0) break;
		p = T->kid[cc > 0][p];
	}
	for (i = 0; i < len[0] && i < len[1]; ++i)
		if (way[0][i] == way[1][i] && c[0][i] <= 0 && c[1][i] >= 0) break;
	if (i == len[0] || i == len[1]) return NIL;
	lca = i, mn = way[0][lca];
	for (i = lca + 1; i < len[0]; ++i)
		if (c[0][i] <= 0) {
			int32_t q = way[0][i], r = T->kid[1][q];
			if (PRI_LT(T, q, mn)) mn = q;
			if (r != NIL && PRI_LT(T, T->best[r], mn)) mn = T->best[r];
		}
	for (i = lca + 1; i < len[1]; ++i)
		if (c[1][i] >= 0) {
			int32_t q = way[1][i], l = T->kid[0][q];
			if (PRI_LT(T, q, mn)) mn = q;
			if (l != NIL && PRI_LT(T, T->best[l], mn)) mn = T->best[l];
		}
	return mn;
}

/* krmq.h:96-108, lower bound only: the largest key <= (ky, ki) */
static int32_t tree_floor(const tree_t *T, int32_t ky, int64_t ki)
{
	int32_t p = T->root, l = NIL;
	while (p != NIL) {
		int c = key_cmp(ky, ki, T, p);
		if (c < 0) p = T->kid[0][p];
		else if (c > 0) l = p, p = T->kid[1][p];
		else return p;
	}
	return l;
}

/* in-order predecessor walk with an explicit root path (krmq.h:345-385 with dir = 0) */
typedef struct { int32_t stack[DEPTH]; int top; } walk_t;
static void walk_to(const tree_t *T, int32_t x, walk_t *w)
{
	int32_t p = T->root;
	w->top = -1;
	while (p != NIL) {
		int c = key_cmp(T->y[x], x, T, p);
		w->stack[++w->top] = p;
		if (c == 0) break;
		p = T->kid[c > 0][p];
	}
}
static int walk_prev(const tree_t *T, walk_t *w)
{
	int32_t p;
	if (w->top < 0) return 0;
	p = T->kid[0][w->stack[w->top]];
	if (p != NIL) {
		for (; p != NIL; p = T->kid[1][p]) w->stack[++w->top] = p;
		return 1;
	}
	do { p = w->stack[w->top--]; } while (w->top >= 0 && p == T->kid[0][w->stack[w->top]]);
	return w->top < 0? 0 : 1;
}

static int32_t simple_score(const mm2o_128_t *ai, const mm2o_128_t *aj, float pen_gap, float pen_skip, int32_t *exact, int32_t *width) /* lchain.c:232-248 */
{
	int32_t dq = (int32_t)ai->y - (int32_t)aj->y, dr, dd, dg, span_j, sc;
	dr = (int32_t)(ai->x - aj->x);
	*width = dd = dr > dq? dr - dq : dq - dr;
	dg = dr < dq? dr : dq;
	span_j = (int32_t)(aj->y >> 32 & 0xff);
	sc = span_j < dg? span_j : dg;
	if (exact) *exact = (dd == 0 && dg <= span_j);
	if (dd || dq > span_j) {
		float lin = pen_gap * (float)dd + pen_skip * (float)dg;
		float lg = dd >= 1? mm2o_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)(lin + .5f * lg);
	}
	return sc;
}

int mm2o_lchain_rmq(int max_dist, int max_dist_inner, int bw, int max_chn_skip, int cap_rmq_size, int min_cnt, int min_sc,
					float pen_gap, float pen_skip, int64_t n, const mm2o_128_t *a,
					uint64_t **u_out, mm2o_128_t **a_out, int *n_a_out)
{
	int32_t *f, *t, *v, n_u, n_v, max_drop = bw;
	int64_t *p, i, i0, st = 0, st_in = 0;
	uint64_t *u;
	tree_t T0, T1;
	*u_out = 0, *a_out = 0, *n_a_out = 0;
	if (n == 0 || a == 0) return 0;
	if (max_dist < bw) max_dist = bw;
	if (max_dist_inner < 0) max_dist_inner = 0;
	if (max_dist_inner > max_dist) max_dist_inner = max_dist;
	p = (int64_t*)malloc(n * 8);
	f = (int32_t*)malloc(n * 4), v = (int32_t*)malloc(n * 4), t = (int32_t*)calloc(n, 4);
	tree_init(&T0, n); tree_init(&T1, n);
	for (i = i0 = 0; i < n; ++i) { /* lchain.c:276-356 */
		int64_t max_j = -1;
		int32_t span = (int32_t)(a[i].y >> 32 & 0xff), max_f = span, yi = (int32_t)a[i].y, q;
		if (i0 < i && a[i0].x != a[i].x) { /* anchors with a smaller x become available */
			int64_t j;
			for (j = i0; j < i; ++j) {
				double pri = -(f[j] + 0.5 * pen_gap * ((int32_t)a[j].x + (int32_t)a[j].y));
				T0.y[j] = (int32_t)a[j].y, T0.pri[j] = pri;
				tree_insert(&T0, (int32_t)j);
				if (max_dist_inner > 0) { T1.y[j] = (int32_t)a[j].y, T1.pri[j] = pri; tree_insert(&T1, (int32_t)j); }
			}
			i0 = i;
		}
		while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + max_dist || (int32_t)(T0.root == NIL? 0 : T0.size[T0.root]) > cap_rmq_size)) {
			if (tree_find(&T0, (int32_t)a[st].y, st) != NIL) tree_erase(&T0, (int32_t)a[st].y, st);
			++st;
		}
		if (max_dist_inner > 0)
			while (st_in < i && (a[i].x >> 32 != a[st_in].x >> 32 || a[i].x > a[st_in].x + max_dist_inner || (int32_t)(T1.root == NIL? 0 : T1.size[T1.root]) > cap_rmq_size)) {
				if (tree_find(&T1, (int32_t)a[st_in].y, st_in) != NIL) tree_erase(&T1, (int32_t)a[st_in].y, st_in);
				++st_in;
			}
		q = tree_rmq(&T0, yi - max_dist, INT32_MAX, yi, 0);
		if (q != NIL) {
			int32_t sc, exact, width, n_skip = 0;
			int64_t j = q;
			sc = f[j] + simple_score(&a[i], &a[j], pen_gap, pen_skip, &exact, &width);
			if (width <= bw && sc > max_f) max_f = sc, max_j = j;
			if (!exact && T1.root != NIL && yi > 0) {
				int32_t lo = tree_floor(&T1, yi - 1, n);
				if (lo != NIL) {
					walk_t w;
					walk_to(&T1, lo, &w);
					while (w.top >= 0) {
						int32_t qq = w.stack[w.top];
						if (T1.y[qq] < yi - max_dist_inner) break;
						j = qq;
						sc = f[j] + simple_score(&a[i], &a[j], pen_gap, pen_skip, 0, &width);
						if (width <= bw) {
							if (sc > max_f) {
								max_f = sc, max_j = j;
								if (n_skip > 0) --n_skip;
							} else if (t[j] == (int32_t)i) {
								if (++n_skip > max_chn_skip) break;
							}
							if (p[j] >= 0) t[p[j]] = (int32_t)i;
						}
						if (!walk_prev(&T1, &w)) break;
					}
				}
			}
		}
		f[i] = max_f, p[i] = max_j;
		v[i] = max_j >= 0 && v[max_j] > max_f? v[max_j] : max_f;
	}
	tree_free(&T0); tree_free(&T1);
	u = mm2o_chain_backtrack(n, f, p, v, t, min_cnt, min_sc, max_drop, &n_u, &n_v);
	free(p); free(f); free(t);
	if (n_u == 0) { free(v); free(u); return 0; }
	*a_out = mm2o_compact(n_u, u, n_v, v, a);
	*u_out = u, *n_a_out = n_v;
	free(v);
	return n_u;
}
