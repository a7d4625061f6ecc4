/* oracle/ref_shim_align.c -- TEST INFRASTRUCTURE ONLY. Reaches the reference's static per-hit tail (mm_append_cigar, mm_fix_cigar,
 * mm_update_extra: align.c:105-334) by compiling the unmodified /root/reference/align.c into this translation unit. Built as its own
 * shared object (oracle/_ref/libminimap2_refalign.so: the reference objects without align.o + this file), never linked into the product. */
#include "align.c"

/* pieces: n_pieces CIGARs back to back in ops[] (piece_len[i] operations each), appended in order with mm_append_cigar; then
 * mm_update_extra (which calls mm_fix_cigar) on the query / target pieces. coor[4] = qs, qe, rs, re (in: before; out: after the
 * leading-I/D shift), out[6] = n_cigar, blen, mlen, n_ambi, dp_max, is_spliced; cig_out receives the final operations. */
void refshim_hit_tail(int n_pieces, const uint32_t *piece_len, const uint32_t *ops, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat,
					  int q, int e, int rev, int32_t *coor, int32_t *out, uint32_t *cig_out)
{
	mm_reg1_t r;
	int i;
	memset(&r, 0, sizeof(r));
	r.rev = rev, r.qs = coor[0], r.qe = coor[1], r.rs = coor[2], r.re = coor[3];
	for (i = 0; i < n_pieces; ++i) {
		if (piece_len[i] > 0) mm_append_cigar(&r, piece_len[i], (uint32_t*)ops);
		ops += piece_len[i];
	}
	out[0] = out[1] = out[2] = out[3] = out[4] = out[5] = 0;
	if (r.p) {
		mm_update_extra(&r, qseq, tseq, mat, (int8_t)q, (int8_t)e, 0, 1);
		out[0] = r.p->n_cigar, out[1] = r.blen, out[2] = r.mlen, out[3] = r.p->n_ambi, out[4] = r.p->dp_max, out[5] = r.is_spliced;
		memcpy(cig_out, r.p->cigar, r.p->n_cigar * 4);
		free(r.p);
	}
	coor[0] = r.qs, coor[1] = r.qe, coor[2] = r.rs, coor[3] = r.re;
}
