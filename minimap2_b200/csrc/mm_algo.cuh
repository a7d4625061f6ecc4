// minimap2_b200/csrc/mm_algo.cuh -- small exact-semantics building blocks shared by device kernels and host
// orchestration (__host__ __device__). Everything here must reproduce the reference bit-for-bit; each block cites
// the reference lines that define the behaviour. Compiled with --fmad=false (float chain penalties, lchain.c:127-135).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef MM_HD
#define MM_HD __host__ __device__ __forceinline__
#endif

struct m128 { uint64_t x, y; };

#define MMX_SEED_LONG_JOIN  (1ULL<<40)  // mmpriv.h:19-25
#define MMX_SEED_IGNORE     (1ULL<<41)
#define MMX_SEED_TANDEM     (1ULL<<42)
#define MMX_SEED_SELF       (1ULL<<43)
#define MMX_SEED_SEG_SHIFT  48
#define MMX_SEED_SEG_MASK   (0xffULL<<MMX_SEED_SEG_SHIFT)

// invertible integer hash on 2k bits (sketch.c:28-38)
MM_HD uint64_t mmx_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

// 64-bit mix used for hit tie-breaking (hit.c:39-49)
MM_HD uint64_t mmx_hash64_full(uint64_t key)
{
	key = (~key + (key << 21));
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8));
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4));
	key = key ^ key >> 28;
	key = (key + (key << 31));
	return key;
}

// ---------------------------------------------------------------------------------------------------------
// Exact emulation of radix_sort_128x / radix_sort_64 (ksort.h:98-151): in-place MSD American-flag sort, 8 bits per
// level starting at byte 7, ranges of <=64 insertion-sorted. Not stable: the permutation of equal keys is defined by
// the cycle-leader walk, which is reproduced step by step. Iterative (explicit range stack) so that it can run as one
// device thread per array. `stk` must hold 3*(n/65*8+8) int32 entries... callers size it with mmx_rs_stack_len(n).
// ---------------------------------------------------------------------------------------------------------
MM_HD int64_t mmx_rs_stack_len(int64_t n) { return 3 * ((n / 65 + 1) * 8 + 8); }

template<class T, class KeyF>
MM_HD void mmx_rs_insertion(T *a, int64_t beg, int64_t end, KeyF key)
{
	for (int64_t i = beg + 1; i < end; ++i)
		if (key(a[i]) < key(a[i - 1])) {
			T tmp = a[i];
			int64_t j;
			for (j = i; j > beg && key(tmp) < key(a[j - 1]); --j) a[j] = a[j - 1];
			a[j] = tmp;
		}
}

template<class T, class KeyF>
__host__ __device__ inline void mmx_rs_sort(T *a, int64_t n, int32_t *stk, KeyF key)
{
	if (n <= 64) { mmx_rs_insertion(a, 0, n, key); return; }
	uint32_t head[256], tail[256];
	int sp = 0;
	stk[sp++] = 0, stk[sp++] = (int32_t)n, stk[sp++] = 56;
	while (sp > 0) {
		int shift = stk[--sp]; const int64_t end = stk[--sp], beg = stk[--sp];
		{ // A level on which every key of the range has the same digit leaves the range untouched (each element is already in its
		  // bucket: ksort.h:126-138 only advances the bucket head) and hands the whole range to the next level. Jump straight to
		  // the most significant byte in which the keys differ; identical keys need no work at all. Typical: chain scores (< 2^20)
		  // sorted as 64-bit keys skip five of eight levels.
			const uint64_t k0 = key(a[beg]);
			uint64_t diff = 0;
			for (int64_t i = beg + 1; i < end; ++i) diff |= key(a[i]) ^ k0;
			if (diff == 0) continue;
			int hb = 7;
			while (!(diff >> (hb * 8) & 0xff)) --hb;
			if (hb * 8 < shift) shift = hb * 8;
		}
		// count
		for (int b = 0; b < 256; ++b) head[b] = 0;
		for (int64_t i = beg; i < end; ++i) ++head[key(a[i]) >> shift & 0xff];
		uint32_t run = (uint32_t)beg;
		for (int b = 0; b < 256; ++b) { uint32_t c = head[b]; head[b] = run; run += c; tail[b] = run; }
		// permute (ksort.h:126-138)
		for (int b = 0; b < 256;) {
			if (head[b] != tail[b]) {
				int l = (int)(key(a[head[b]]) >> shift & 0xff);
				if (l != b) {
					T tmp = a[head[b]], sw;
					do {
						sw = tmp; tmp = a[head[l]]; a[head[l]++] = sw;
						l = (int)(key(tmp) >> shift & 0xff);
					} while (l != b);
					a[head[b]++] = tmp;
				} else ++head[b];
			} else ++b;
		}
		if (shift) {
			const int ns = shift > 8? shift - 8 : 0;
			uint32_t bb = (uint32_t)beg;
			for (int b = 0; b < 256; ++b) {
				const uint32_t be = tail[b];
				if (be - bb > 64) stk[sp++] = (int32_t)bb, stk[sp++] = (int32_t)be, stk[sp++] = ns;
				else if (be - bb > 1) mmx_rs_insertion(a, (int64_t)bb, (int64_t)be, key);
				bb = be;
			}
		}
	}
}

struct KeyX128 { MM_HD uint64_t operator()(const m128 &v) const { return v.x; } };
struct KeyU64  { MM_HD uint64_t operator()(const uint64_t &v) const { return v; } };

// ---------------------------------------------------------------------------------------------------------
// chaining score (lchain.c:113-138) and mg_log2 (mmpriv.h:139-147)
// ---------------------------------------------------------------------------------------------------------
MM_HD float mmx_log2(float x)
{
	union { float f; uint32_t i; } z; z.f = x;
	float r = (float)((int)((z.i >> 23) & 255) - 128);
	z.i &= ~(255U << 23);
	z.i += 127U << 23;
	r += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return r;
}

MM_HD int32_t mmx_comput_sc(const m128 &ai, const m128 &aj, int32_t max_dist_x, int32_t max_dist_y, int32_t bw,
							float pen_gap, float pen_skip, int is_cdna, int n_seg)
{
	int32_t dq = (int32_t)ai.y - (int32_t)aj.y, dr, dd, dg, span_j, sc;
	int32_t si = (int32_t)((ai.y & MMX_SEED_SEG_MASK) >> MMX_SEED_SEG_SHIFT), sj = (int32_t)((aj.y & MMX_SEED_SEG_MASK) >> MMX_SEED_SEG_SHIFT);
	if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
	dr = (int32_t)(ai.x - aj.x);
	if (si == sj && (dr == 0 || dq > max_dist_y)) return INT32_MIN;
	dd = dr > dq? dr - dq : dq - dr;
	if (si == sj && dd > bw) return INT32_MIN;
	if (n_seg > 1 && !is_cdna && si == sj && dr > max_dist_y) return INT32_MIN;
	dg = dr < dq? dr : dq;
	span_j = (int32_t)(aj.y >> 32 & 0xff);
	sc = span_j < dg? span_j : dg;
	if (dd || dg > span_j) {
		float lin = pen_gap * (float)dd + pen_skip * (float)dg;
		float lg = dd >= 1? mmx_log2((float)(dd + 1)) : 0.0f;
		if (is_cdna || si != sj) {
			if (si != sj && dr == 0) ++sc;
			else if (dr > dq || si != sj) sc -= (int)(lin < lg? lin : lg);
			else sc -= (int)(lin + .5f * lg);
		} else sc -= (int)(lin + .5f * lg);
	}
	return sc;
}

// rmq variant (lchain.c:232-248)
MM_HD int32_t mmx_comput_sc_simple(const m128 &ai, const m128 &aj, float pen_gap, float pen_skip, int32_t *exact, int32_t *width)
{
	int32_t dq = (int32_t)ai.y - (int32_t)aj.y, dr, dd, dg, span_j, sc;
	dr = (int32_t)(ai.x - aj.x);
	*width = dd = dr > dq? dr - dq : dq - dr;
	dg = dr < dq? dr : dq;
	span_j = (int32_t)(aj.y >> 32 & 0xff);
	sc = span_j < dg? span_j : dg;
	if (exact) *exact = (dd == 0 && dg <= span_j);
	if (dd || dq > span_j) {
		float lin = pen_gap * (float)dd + pen_skip * (float)dg;
		float lg = dd >= 1? mmx_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)(lin + .5f * lg);
	}
	return sc;
}

// 4-bit packed reference access (mmpriv.h:34-35)
MM_HD uint32_t mmx_seq4_get(const uint32_t *S, uint64_t i) { return S[i >> 3] >> ((i & 7) << 2) & 0xf; }

MM_HD uint8_t mmx_nt4(uint8_t c) // sketch.c:9-26
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	default: return c < 4? c : 4;
	}
}
