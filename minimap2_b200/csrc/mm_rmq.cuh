// minimap2_b200/csrc/mm_rmq.cuh -- index-based AVL tree with subtree-minimum augmentation, reproducing the exact shape
// and tie behaviour of the reference's krmq.h instantiation in lchain.c:219-230 (key = (y, i), priority = pri).
// mg_lchain_rmq's result depends on the tree SHAPE (krmq_rmq returns the first minimum met while walking the two root
// paths, krmq.h:110-157), so insertion/deletion/rotation follow krmq.h:160-330 step by step. Nodes are array slots
// (node id == anchor index; slot n is the "fake" super-root used by erase), links are int32 indices, -1 == NULL.
#pragma once
#include "mm_algo.cuh"

#define RMQ_MAX_DEPTH 64
#define RMQ_NIL (-1)

struct RmqTree {
	int32_t *c[2];    // children
	int32_t *s;       // slot of the minimum-priority node in the subtree
	int8_t *bal;
	uint32_t *size;
	int32_t *y;       // key part 1 (key part 2 is the slot id itself, i.e. the anchor index)
	double *pri;
	int32_t root;
	int32_t n;        // slot n = fake node
};

MM_HD int rmq_cmp_key(int32_t ky, int64_t ki, const RmqTree &T, int32_t b) // lc_elem_cmp(a,b) with a given by (ky,ki)
{
	const int32_t by = T.y[b];
	return ky < by? -1 : ky > by? 1 : (ki > (int64_t)b) - (ki < (int64_t)b);
}
MM_HD bool rmq_lt2(const RmqTree &T, int32_t a, int32_t b) { return T.pri[a] < T.pri[b]; }
MM_HD uint32_t rmq_size_child(const RmqTree &T, int32_t q, int i) { return T.c[i][q] >= 0? T.size[T.c[i][q]] : 0; }

MM_HD int32_t rmq_find(const RmqTree &T, int32_t ky, int64_t ki) // krmq.h:83-95
{
	int32_t p = T.root;
	while (p >= 0) {
		int cmp = rmq_cmp_key(ky, ki, T, p);
		if (cmp < 0) p = T.c[0][p];
		else if (cmp > 0) p = T.c[1][p];
		else break;
	}
	return p;
}

MM_HD int32_t rmq_interval(const RmqTree &T, int32_t ky, int64_t ki, int32_t *lower, int32_t *upper) // krmq.h:96-108
{
	int32_t p = T.root, l = RMQ_NIL, u = RMQ_NIL;
	while (p >= 0) {
		int cmp = rmq_cmp_key(ky, ki, T, p);
		if (cmp < 0) u = p, p = T.c[0][p];
		else if (cmp > 0) l = p, p = T.c[1][p];
		else { l = u = p; break; }
	}
	*lower = l, *upper = u;
	return p;
}

// closed-interval range-minimum query (krmq.h:110-151)
__host__ __device__ inline int32_t rmq_query(const RmqTree &T, int32_t lo_y, int64_t lo_i, int32_t up_y, int64_t up_i)
{
	int32_t p = T.root, path[2][RMQ_MAX_DEPTH], mn;
	int plen[2] = {0, 0}, i, cmp, lca;
	int8_t pcmp[2][RMQ_MAX_DEPTH];
	if (T.root < 0) return RMQ_NIL;
	while (p >= 0) {
		cmp = rmq_cmp_key(lo_y, lo_i, T, p);
		path[0][plen[0]] = p, pcmp[0][plen[0]++] = (int8_t)cmp;
		if (cmp < 0) p = T.c[0][p];
		else if (cmp > 0) p = T.c[1][p];
		else break;
	}
	p = T.root;
	while (p >= 0) {
		cmp = rmq_cmp_key(up_y, up_i, T, p);
		path[1][plen[1]] = p, pcmp[1][plen[1]++] = (int8_t)cmp;
		if (cmp < 0) p = T.c[0][p];
		else if (cmp > 0) p = T.c[1][p];
		else break;
	}
	for (i = 0; i < plen[0] && i < plen[1]; ++i)
		if (path[0][i] == path[1][i] && pcmp[0][i] <= 0 && pcmp[1][i] >= 0) break;
	if (i == plen[0] || i == plen[1]) return RMQ_NIL;
	lca = i, mn = path[0][lca];
	for (i = lca + 1; i < plen[0]; ++i) {
		if (pcmp[0][i] <= 0) {
			const int32_t q = path[0][i];
			if (rmq_lt2(T, q, mn)) mn = q;
			if (T.c[1][q] >= 0 && rmq_lt2(T, T.s[T.c[1][q]], mn)) mn = T.s[T.c[1][q]];
		}
	}
	for (i = lca + 1; i < plen[1]; ++i) {
		if (pcmp[1][i] >= 0) {
			const int32_t q = path[1][i];
			if (rmq_lt2(T, q, mn)) mn = q;
			if (T.c[0][q] >= 0 && rmq_lt2(T, T.s[T.c[0][q]], mn)) mn = T.s[T.c[0][q]];
		}
	}
	return mn;
}

MM_HD void rmq_update_min(RmqTree &T, int32_t p, int32_t q, int32_t r) // krmq.h:155-158
{
	T.s[p] = (q < 0 || rmq_lt2(T, p, T.s[q]))? p : T.s[q];
	T.s[p] = (r < 0 || rmq_lt2(T, T.s[p], T.s[r]))? T.s[p] : T.s[r];
}

MM_HD int32_t rmq_rotate1(RmqTree &T, int32_t p, int dir) // krmq.h:160-172
{
	const int opp = 1 - dir;
	const int32_t q = T.c[opp][p], s = T.s[p];
	const uint32_t size_p = T.size[p];
	T.size[p] -= T.size[q] - rmq_size_child(T, q, dir);
	T.size[q] = size_p;
	rmq_update_min(T, p, T.c[dir][p], T.c[dir][q]);
	T.s[q] = s;
	T.c[opp][p] = T.c[dir][q];
	T.c[dir][q] = p;
	return q;
}

MM_HD int32_t rmq_rotate2(RmqTree &T, int32_t p, int dir) // krmq.h:174-195
{
	const int opp = 1 - dir;
	const int32_t q = T.c[opp][p], r = T.c[dir][q], s = T.s[p];
	const uint32_t size_x_dir = rmq_size_child(T, r, dir);
	int b1;
	T.size[r] = T.size[p];
	T.size[p] -= T.size[q] - size_x_dir;
	T.size[q] -= size_x_dir + 1;
	rmq_update_min(T, p, T.c[dir][p], T.c[dir][r]);
	rmq_update_min(T, q, T.c[opp][q], T.c[opp][r]);
	T.s[r] = s;
	T.c[opp][p] = T.c[dir][r];
	T.c[dir][r] = p;
	T.c[dir][q] = T.c[opp][r];
	T.c[opp][r] = q;
	b1 = dir == 0? +1 : -1;
	if (T.bal[r] == b1) T.bal[q] = 0, T.bal[p] = (int8_t)-b1;
	else if (T.bal[r] == 0) T.bal[q] = T.bal[p] = 0;
	else T.bal[q] = (int8_t)b1, T.bal[p] = 0;
	T.bal[r] = 0;
	return r;
}

// insert slot x (its y/pri must be set). Returns x, or the existing equal node (krmq.h:197-244)
__host__ __device__ inline int32_t rmq_insert(RmqTree &T, int32_t x)
{
	uint8_t stack[RMQ_MAX_DEPTH];
	int32_t path[RMQ_MAX_DEPTH];
	int32_t bp = T.root, bq = RMQ_NIL, p, q, r = RMQ_NIL;
	int i, which = 0, top = 0, b1, path_len = 0;
	for (p = bp, q = bq; p >= 0; q = p, p = T.c[which][p]) {
		int cmp = rmq_cmp_key(T.y[x], (int64_t)x, T, p);
		if (cmp == 0) return p;
		if (T.bal[p] != 0) bq = q, bp = p, top = 0;
		stack[top++] = (uint8_t)(which = (cmp > 0));
		path[path_len++] = p;
	}
	T.bal[x] = 0, T.size[x] = 1, T.c[0][x] = T.c[1][x] = RMQ_NIL, T.s[x] = x;
	if (q < 0) T.root = x;
	else T.c[which][q] = x;
	if (bp < 0) return x;
	for (i = 0; i < path_len; ++i) ++T.size[path[i]];
	for (i = path_len - 1; i >= 0; --i) {
		rmq_update_min(T, path[i], T.c[0][path[i]], T.c[1][path[i]]);
		if (T.s[path[i]] != x) break;
	}
	for (p = bp, top = 0; p != x; p = T.c[stack[top]][p], ++top) {
		if (stack[top] == 0) --T.bal[p];
		else ++T.bal[p];
	}
	if (T.bal[bp] > -2 && T.bal[bp] < 2) return x;
	which = (T.bal[bp] < 0);
	b1 = which == 0? +1 : -1;
	q = T.c[1 - which][bp];
	if (T.bal[q] == b1) {
		r = rmq_rotate1(T, bp, which);
		T.bal[q] = T.bal[bp] = 0;
	} else r = rmq_rotate2(T, bp, which);
	if (bq < 0) T.root = r;
	else T.c[bp != T.c[0][bq]][bq] = r;
	return x;
}

// erase the node equal to key (ky,ki); returns the removed slot or -1 (krmq.h:246-330). Slot T.n plays `fake`.
__host__ __device__ inline int32_t rmq_erase(RmqTree &T, int32_t ky, int64_t ki)
{
	int32_t p, path[RMQ_MAX_DEPTH];
	uint8_t dir[RMQ_MAX_DEPTH];
	int i, d = 0, cmp;
	const int32_t fake = T.n;
	if (T.root < 0) return RMQ_NIL;
	// fake = *root; fake.p[0] = root; fake.p[1] = 0
	T.y[fake] = T.y[T.root], T.pri[fake] = T.pri[T.root], T.s[fake] = T.s[T.root], T.bal[fake] = T.bal[T.root], T.size[fake] = T.size[T.root];
	T.c[0][fake] = T.root, T.c[1][fake] = RMQ_NIL;
	for (cmp = -1, p = fake; cmp; cmp = rmq_cmp_key(ky, ki, T, p)) {
		int which = (cmp > 0);
		dir[d] = (uint8_t)which;
		path[d++] = p;
		p = T.c[which][p];
		if (p < 0) return RMQ_NIL;
	}
	for (i = 1; i < d; ++i) --T.size[path[i]];
	if (T.c[1][p] < 0) {
		T.c[dir[d-1]][path[d-1]] = T.c[0][p];
	} else {
		int32_t q = T.c[1][p];
		if (T.c[0][q] < 0) {
			T.c[0][q] = T.c[0][p];
			T.bal[q] = T.bal[p];
			T.c[dir[d-1]][path[d-1]] = q;
			path[d] = q, dir[d++] = 1;
			T.size[q] = T.size[p] - 1;
		} else {
			int32_t r;
			int e = d++;
			for (;;) {
				dir[d] = 0;
				path[d++] = q;
				r = T.c[0][q];
				if (T.c[0][r] < 0) break;
				q = r;
			}
			T.c[0][r] = T.c[0][p];
			T.c[0][q] = T.c[1][r];
			T.c[1][r] = T.c[1][p];
			T.bal[r] = T.bal[p];
			T.c[dir[e-1]][path[e-1]] = r;
			path[e] = r, dir[e] = 1;
			for (i = e + 1; i < d; ++i) --T.size[path[i]];
			T.size[r] = T.size[p] - 1;
		}
	}
	for (i = d - 1; i >= 0; --i) rmq_update_min(T, path[i], T.c[0][path[i]], T.c[1][path[i]]);
	while (--d > 0) {
		int32_t q = path[d];
		int which, other, b1 = 1, b2 = 2;
		which = dir[d], other = 1 - which;
		if (which) b1 = -b1, b2 = -b2;
		T.bal[q] += (int8_t)b1;
		if (T.bal[q] == b1) break;
		else if (T.bal[q] == b2) {
			int32_t r = T.c[other][q];
			if (T.bal[r] == -b1) {
				T.c[dir[d-1]][path[d-1]] = rmq_rotate2(T, q, which);
			} else {
				T.c[dir[d-1]][path[d-1]] = rmq_rotate1(T, q, which);
				if (T.bal[r] == 0) {
					T.bal[r] = (int8_t)-b1;
					T.bal[q] = (int8_t)b1;
					break;
				} else T.bal[r] = T.bal[q] = 0;
			}
		}
	}
	T.root = T.c[0][fake];
	return p;
}

// iterator (krmq.h:345-385)
struct RmqItr { int32_t stack[RMQ_MAX_DEPTH]; int top; };

MM_HD int rmq_itr_find(const RmqTree &T, int32_t ky, int64_t ki, RmqItr *itr)
{
	int32_t p = T.root;
	itr->top = -1;
	while (p >= 0) {
		itr->stack[++itr->top] = p;
		int cmp = rmq_cmp_key(ky, ki, T, p);
		if (cmp < 0) p = T.c[0][p];
		else if (cmp > 0) p = T.c[1][p];
		else break;
	}
	return p >= 0? 1 : 0;
}
MM_HD int32_t rmq_itr_at(const RmqItr *itr) { return itr->top < 0? RMQ_NIL : itr->stack[itr->top]; }
MM_HD int rmq_itr_next_bidir(const RmqTree &T, RmqItr *itr, int dir)
{
	int32_t p;
	if (itr->top < 0) return 0;
	dir = !!dir;
	p = T.c[dir][itr->stack[itr->top]];
	if (p >= 0) {
		for (; p >= 0; p = T.c[!dir][p]) itr->stack[++itr->top] = p;
		return 1;
	} else {
		do { p = itr->stack[itr->top--]; } while (itr->top >= 0 && p == T.c[dir][itr->stack[itr->top]]);
		return itr->top < 0? 0 : 1;
	}
}
