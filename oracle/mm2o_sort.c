/* oracle/mm2o_sort.c -- TEST INFRASTRUCTURE ONLY (see mm2o.h).
 * Restates radix_sort_128x / radix_sort_64 (ksort.h:98-151, misc.c:155-159): in-place MSD American-flag sort,
 * 8 bits per level from byte 7, insertion sort for ranges of <=64. NOT stable; the exact permutation of equal keys is
 * part of the contract (it decides chain peel order in lchain.c:38-71), so the cycle-leader walk is restated exactly. */
#include <string.h>
#include "mm2o.h"

#define MIN_RANGE 64

#define DEF_SORT(NAME, T, KEY) \
static void ins_##NAME(T *beg, T *end) { \
	T *i; \
	for (i = beg + 1; i < end; ++i) \
		if (KEY(*i) < KEY(*(i - 1))) { \
			T *j, tmp = *i; \
			for (j = i; j > beg && KEY(tmp) < KEY(*(j - 1)); --j) *j = *(j - 1); \
			*j = tmp; \
		} \
} \
static void flag_##NAME(T *beg, T *end, int shift) { \
	T *head[256], *tail[256]; /* head: next unplaced slot of each bucket; tail: bucket end */ \
	size_t cnt[256]; \
	int b; T *i; \
	memset(cnt, 0, sizeof(cnt)); \
	for (i = beg; i != end; ++i) ++cnt[KEY(*i) >> shift & 0xff]; \
	for (b = 0, i = beg; b < 256; ++b) head[b] = i, i += cnt[b], tail[b] = i; \
	for (b = 0; b < 256;) { \
		if (head[b] != tail[b]) { \
			int l = (int)(KEY(*head[b]) >> shift & 0xff); \
			if (l != b) { /* chase the cycle starting from the first unplaced element of bucket b */ \
				T tmp = *head[b], sw; \
				do { sw = tmp; tmp = *head[l]; *head[l]++ = sw; l = (int)(KEY(tmp) >> shift & 0xff); } while (l != b); \
				*head[b]++ = tmp; \
			} else ++head[b]; \
		} else ++b; \
	} \
	if (shift) { \
		int ns = shift > 8? shift - 8 : 0; \
		T *bb = beg; \
		for (b = 0; b < 256; ++b) { \
			T *be = tail[b]; \
			if (be - bb > MIN_RANGE) flag_##NAME(bb, be, ns); \
			else if (be - bb > 1) ins_##NAME(bb, be); \
			bb = be; \
		} \
	} \
} \
void mm2o_radix_sort_##NAME(T *beg, T *end) { \
	if (end - beg <= MIN_RANGE) ins_##NAME(beg, end); \
	else flag_##NAME(beg, end, 56); \
}

#define KEY128(a) ((a).x)
#define KEY64(a) (a)
DEF_SORT(128x, mm2o_128_t, KEY128)
DEF_SORT(64, uint64_t, KEY64)
