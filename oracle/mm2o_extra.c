/* oracle/mm2o_extra.c -- TEST INFRASTRUCTURE ONLY (the checker for kernel K4, never linked into the product).
 * Plain-C restatement of the per-hit tail of the alignment driver: stitching the ksw2 pieces of a hit into one CIGAR
 * (mm_append_cigar, align.c:317-334), left-aligning gaps / collapsing I-D runs / dropping a leading gap (mm_fix_cigar,
 * align.c:105-181) and the per-base statistics with the clamped running score (mm_update_extra, align.c:254-303; the gap
 * term uses mg_log2, mmpriv.h:139-147). Arithmetic is the reference's: the running score is a double.
 * Pinned against the reference's own static functions through oracle/_ref/libminimap2_refalign.so (tests/test_oracle_vs_ref.py)
 * and by the committed vectors in tests/golden/vectors_tail.npz. */
#include <stdint.h>
#include <string.h>
#include "mm2o.h"

float mm2o_log2(float x);

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3 };

/* coor[4] = qs, qe, rs, re (updated when a leading gap is dropped); out[6] = n_cigar, blen, mlen, n_ambi, dp_max, is_spliced.
 * cig must have room for the sum of the piece lengths. qseq / tseq: the hit's query piece on its strand and target piece (nt4). */
void mm2o_hit_tail(int n_pieces, const uint32_t *piece_len, const uint32_t *ops, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat,
				   int q, int e, int rev, int32_t *coor, int32_t *out, uint32_t *cig)
{
	uint32_t n = 0, k;
	int i, shrink = 0;
	int32_t qo = 0, to = 0, blen = 0, mlen = 0, n_ambi = 0, spliced = 0;
	double s = 0.0, best = 0.0;
	/* stitching: only the first operation of a piece can merge, into the last one so far (align.c:323-333) */
	for (i = 0; i < n_pieces; ops += piece_len[i], ++i) {
		uint32_t m = piece_len[i], j = 0;
		if (m == 0) continue;
		if (n > 0 && (cig[n-1] & 0xf) == (ops[0] & 0xf)) cig[n-1] += ops[0] >> 4 << 4, j = 1;
		for (; j < m; ++j) cig[n++] = ops[j];
	}
	memset(out, 0, 6 * sizeof(int32_t));
	if (n == 0) return; /* no mm_extra_t: mm_update_extra returns at once (align.c:260) */
	if (n > 1) { /* mm_fix_cigar does nothing on a single operation (align.c:110) */
		for (k = 0; k < n; ++k) { /* align.c:111-132: shift a gap between two matches left while the bases it would swap are equal */
			uint32_t op = cig[k] & 0xf, len = cig[k] >> 4;
			if (len == 0) shrink = 1;
			if (op == OP_M) to += len, qo += len;
			else if (op == OP_I || op == OP_D) {
				if (k > 0 && k + 1 < n && (cig[k-1] & 0xf) == OP_M && (cig[k+1] & 0xf) == OP_M) {
					int32_t l, before = cig[k-1] >> 4;
					const uint8_t *sq = op == OP_I? qseq + qo : tseq + to; /* the gapped sequence at the gap start */
					for (l = 0; l < before; ++l) if (sq[-1 - l] != sq[(int32_t)len - 1 - l]) break;
					if (l > 0) cig[k-1] -= (uint32_t)l << 4, cig[k+1] += (uint32_t)l << 4, qo -= l, to -= l;
					if (l == before) shrink = 1;
				}
				if (op == OP_I) qo += len; else to += len;
			} else if (op == OP_N) to += len;
		}
		for (k = 0; k + 2 < n; ++k) { /* align.c:134-155: a run of adjacent I/D operations becomes one I and one D */
			if ((cig[k] & 0xf) > 0 && (cig[k] & 0xf) + (cig[k+1] & 0xf) == 3) {
				uint32_t l, sum[4] = {0, 0, 0, 0};
				for (l = k; l < n; ++l) {
					uint32_t op = cig[l] & 0xf;
					if (op == OP_I || op == OP_D || cig[l] >> 4 == 0) sum[op & 3] += cig[l] >> 4;
					else break;
				}
				if (sum[OP_I] > 0 && sum[OP_D] > 0 && l - k > 2) {
					cig[k] = sum[OP_I] << 4 | OP_I, cig[k+1] = sum[OP_D] << 4 | OP_D;
					for (k += 2; k < l; ++k) cig[k] &= 0xf;
					shrink = 1;
				}
				k = l;
			}
		}
		if (shrink) { /* align.c:156-169 */
			uint32_t l = 0;
			for (k = 0; k < n; ++k) if (cig[k] >> 4) cig[l++] = cig[k];
			n = l, l = 0;
			for (k = 0; k < n; ++k) {
				if (k + 1 == n || (cig[k] & 0xf) != (cig[k+1] & 0xf)) cig[l++] = cig[k];
				else cig[k+1] += cig[k] >> 4 << 4;
			}
			n = l;
		}
		if ((cig[0] & 0xf) == OP_I || (cig[0] & 0xf) == OP_D) { /* align.c:170-180 */
			int32_t l = cig[0] >> 4;
			if ((cig[0] & 0xf) == OP_I) { if (rev) coor[1] -= l; else coor[0] += l; qseq += l; }
			else coor[2] += l, tseq += l;
			--n;
			memmove(cig, cig + 1, n * 4);
		}
	}
	qo = to = 0;
	for (k = 0; k < n; ++k) { /* align.c:266-297 */
		uint32_t op = cig[k] & 0xf, len = cig[k] >> 4, l;
		if (op == OP_M) {
			int amb = 0, diff = 0;
			for (l = 0; l < len; ++l) {
				int cq = qseq[qo + l], ct = tseq[to + l];
				if (ct > 3 || cq > 3) ++amb;
				else if (ct != cq) ++diff;
				s += mat[ct * 5 + cq];
				if (s < 0) s = 0;
				else if (s > best) best = s;
			}
			blen += len - amb, mlen += len - (amb + diff), n_ambi += amb;
			to += len, qo += len;
		} else if (op == OP_I || op == OP_D) {
			int amb = 0;
			const uint8_t *sq = op == OP_I? qseq + qo : tseq + to;
			for (l = 0; l < len; ++l) if (sq[l] > 3) ++amb;
			blen += len - amb, n_ambi += amb;
			s -= q + (double)e * mm2o_log2(1.0 + len);
			if (s < 0) s = 0;
			if (op == OP_I) qo += len; else to += len;
		} else if (op == OP_N) spliced = 1, to += len;
	}
	out[0] = (int32_t)n, out[1] = blen, out[2] = mlen, out[3] = n_ambi, out[4] = (int32_t)(best + .499), out[5] = spliced;
}
