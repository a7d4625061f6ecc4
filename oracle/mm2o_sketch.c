/* oracle/mm2o_sketch.c -- TEST INFRASTRUCTURE ONLY (see mm2o.h).
 * Restates mm_sketch() (sketch.c:77-143) and hash64() (sketch.c:28-38). Written as an explicit window
 * state machine: W = ring of the last w slot values; cur = current minimum and its slot. */
#include <string.h>
#include "mm2o.h"

static const uint64_t U64MAX = ~(uint64_t)0;

static inline int nt4(unsigned char c) /* sketch.c:9-26 seq_nt4_table */
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	default: return c < 4? c : 4; /* table maps bytes 0..3 to themselves */
	}
}

uint64_t mm2o_hash64(uint64_t key, uint64_t mask) /* sketch.c:28-38 */
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

int mm2o_sketch(const char *seq, int len, int w, int k, uint32_t rid, int is_hpc, mm2o_128_t *out)
{
	uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1, kmer[2] = {0, 0};
	mm2o_128_t W[256], cur = { U64MAX, U64MAX };
	int hp_q[32], hp_front = 0, hp_cnt = 0; /* tiny queue of homopolymer run lengths (sketch.c:40-60) */
	int i, j, l = 0, wpos = 0, cur_pos = 0, span = 0, n = 0;
	if (!(len > 0 && w > 0 && w < 256 && k > 0 && k <= 28)) return -1;
	memset(W, 0xff, sizeof(mm2o_128_t) * w);
	for (i = 0; i < len; ++i) {
		int c = nt4((unsigned char)seq[i]);
		mm2o_128_t info = { U64MAX, U64MAX };
		if (c < 4) {
			int z;
			if (is_hpc) { /* sketch.c:94-104 */
				int run = 1;
				if (i + 1 < len && nt4((unsigned char)seq[i + 1]) == c) {
					for (run = 2; i + run < len; ++run)
						if (nt4((unsigned char)seq[i + run]) != c) break;
					i += run - 1;
				}
				hp_q[(hp_cnt++ + hp_front) & 0x1f] = run;
				span += run;
				if (hp_cnt > k) { span -= hp_q[hp_front++]; hp_front &= 0x1f; --hp_cnt; }
			} else span = l + 1 < k? l + 1 : k;
			kmer[0] = (kmer[0] << 2 | c) & mask;
			kmer[1] = (kmer[1] >> 2) | (3ULL ^ c) << shift1;
			if (kmer[0] == kmer[1]) continue; /* symmetric k-mer: no slot consumed (sketch.c:108) */
			z = kmer[0] < kmer[1]? 0 : 1;
			++l;
			if (l >= k && span < 256) {
				info.x = mm2o_hash64(kmer[z], mask) << 8 | span;
				info.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | z;
			}
		} else l = 0, hp_cnt = hp_front = 0, span = 0;
		W[wpos] = info;
		if (l == w + k - 1 && cur.x != U64MAX) { /* first full window: emit copies of the minimum (sketch.c:117-122) */
			for (j = wpos + 1; j < w; ++j) if (cur.x == W[j].x && W[j].y != cur.y) out[n++] = W[j];
			for (j = 0; j < wpos; ++j)     if (cur.x == W[j].x && W[j].y != cur.y) out[n++] = W[j];
		}
		if (info.x <= cur.x) { /* sketch.c:123-125 */
			if (l >= w + k && cur.x != U64MAX) out[n++] = cur;
			cur = info, cur_pos = wpos;
		} else if (wpos == cur_pos) { /* minimum left the window (sketch.c:126-138) */
			if (l >= w + k - 1 && cur.x != U64MAX) out[n++] = cur;
			for (j = wpos + 1, cur.x = U64MAX; j < w; ++j) if (cur.x >= W[j].x) cur = W[j], cur_pos = j;
			for (j = 0; j <= wpos; ++j)                     if (cur.x >= W[j].x) cur = W[j], cur_pos = j;
			if (l >= w + k - 1 && cur.x != U64MAX) {
				for (j = wpos + 1; j < w; ++j) if (cur.x == W[j].x && cur.y != W[j].y) out[n++] = W[j];
				for (j = 0; j <= wpos; ++j)    if (cur.x == W[j].x && cur.y != W[j].y) out[n++] = W[j];
			}
		}
		if (++wpos == w) wpos = 0;
	}
	if (cur.x != U64MAX) out[n++] = cur;
	return n;
}
