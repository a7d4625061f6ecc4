/* oracle/mm2o_exts2.c -- TEST INFRASTRUCTURE ONLY (see mm2o.h).
 *
 * Scalar restatement of the reference's spliced-alignment kernel ksw_exts2_sse (ksw2_exts2_sse.c:26-465): affine gap
 * (q, e) plus one long "intron" gap state (open q2, free extension) whose opening/closing is scored against the
 * donor/acceptor signals of the target. Like mm2o_extd2, it models one byte lane of the SSE code at a time over the same
 * flat int8 arrays (u|v|x|y|x2|donor|acceptor|s), so the 16-lane rounding of every anti-diagonal is reproduced.
 * There is no band: every anti-diagonal covers [max(0, r-qlen+1), min(r, tlen-1)] (:236-240).
 * Not wired into any product code yet: it is the checker for the next hot-path row (-x splice).
 */
#include <stdlib.h>
#include <string.h>
#include "mm2o.h"

#define I8(v) ((int8_t)(v))

static void s_reset(mm2o_ez_t *ez) /* ksw2.h:164-169 */
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = MM2O_KSW_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
}

static void s_push(mm2o_ez_t *ez, uint32_t op, int len) /* ksw2.h:114-124 */
{
	if (ez->n_cigar == 0 || op != (ez->cigar[ez->n_cigar - 1] & 0xf)) {
		if (ez->n_cigar == ez->m_cigar) {
			ez->m_cigar = ez->m_cigar? ez->m_cigar << 1 : 4;
			ez->cigar = (uint32_t*)realloc(ez->cigar, (size_t)ez->m_cigar << 2);
		}
		ez->cigar[ez->n_cigar++] = (uint32_t)len << 4 | op;
	} else ez->cigar[ez->n_cigar - 1] += (uint32_t)len << 4;
}

/* ksw2.h:130-162 for the rotated layout: state 3 is an intron (N) once min_intron_len > 0 */
static void s_backtrack(mm2o_ez_t *ez, int is_rev, int min_intron_len, const uint8_t *p, const int *off, const int *off_end,
						int n_col, int i0, int j0)
{
	int i = i0, j = j0, state = 0;
	ez->n_cigar = 0;
	while (i >= 0 && j >= 0) {
		int r = i + j, force = -1;
		uint32_t tmp;
		if (i < off[r]) force = 2;
		if (i > off_end[r]) force = 1;
		tmp = force < 0? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7;
		else if (!(tmp >> (state + 2) & 1)) state = 0;
		if (state == 0) state = tmp & 7;
		if (force >= 0) state = force;
		if (state == 0) s_push(ez, 0, 1), --i, --j;
		else if (state == 1 || (state == 3 && min_intron_len <= 0)) s_push(ez, 2, 1), --i;
		else if (state == 3 && min_intron_len > 0) s_push(ez, 3, 1), --i;
		else s_push(ez, 1, 1), --j;
	}
	if (i >= 0) s_push(ez, min_intron_len > 0 && i >= min_intron_len? 3 : 2, i + 1);
	if (j >= 0) s_push(ez, 1, j + 1);
	if (!is_rev) {
		int k;
		for (k = 0; k < ez->n_cigar >> 1; ++k) {
			uint32_t t = ez->cigar[k];
			ez->cigar[k] = ez->cigar[ez->n_cigar - 1 - k], ez->cigar[ez->n_cigar - 1 - k] = t;
		}
	}
}

static int s_zdrop(mm2o_ez_t *ez, int32_t H, int r, int t, int zdrop, int e) /* ksw2.h:171-187 */
{
	if (H > ez->max) {
		ez->max = H, ez->max_t = t, ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql? tl - ql : ql - tl;
		if (zdrop >= 0 && ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

/* donor/acceptor signal arrays (ksw2_exts2_sse.c:111-226). d[t]/a[t] are added when an intron is opened after / closed at
 * target position t; 0 for the canonical signal, -sp[k] for weaker ones, -sp[3] for none. */
static void splice_signals(int tlen, const uint8_t *target, int flag, int noncan, int junc_bonus, int junc_pen, const uint8_t *junc,
						   int8_t *donor, int8_t *acceptor, int tlen16)
{
	static const int sp0[4] = { 8, 15, 21, 30 };
	int sp[4], t;
	if (flag & (MM2O_EZ_SPLICE_FOR | MM2O_EZ_SPLICE_REV)) {
		if (flag & MM2O_EZ_SPLICE_CMPLX) {
			for (t = 0; t < 4; ++t) sp[t] = (int)((double)sp0[t] / 3. + .499);
		} else {
			sp[0] = flag & MM2O_EZ_SPLICE_FLANK? noncan / 2 : 0;
			sp[1] = sp[2] = sp[3] = noncan;
		}
		memset(donor, -sp[3], tlen16), memset(acceptor, -sp[3], tlen16);
		if (!(flag & MM2O_EZ_REV_CIGAR)) {
			for (t = 0; t < tlen - 4; ++t) { /* donor: what follows position t */
				int z = 3;
				const uint8_t c1 = target[t+1], c2 = target[t+2], c3 = target[t+3];
				if (flag & MM2O_EZ_SPLICE_FOR) {
					if (c1 == 2 && c2 == 3) z = c3 == 0 || c3 == 2? -1 : 0;  /* GT[AG] is the best donor */
					else if (c1 == 2 && c2 == 1) z = 1;                       /* GC */
					else if (c1 == 0 && c2 == 3) z = 2;                       /* AT */
				} else if (flag & MM2O_EZ_SPLICE_REV) {
					if (c1 == 1 && c2 == 3) z = c3 == 0 || c3 == 2? -1 : 0;  /* CT: reverse complement of AG */
					else if (c1 == 2 && c2 == 3) z = 2;                       /* GT: reverse complement of AC */
				}
				donor[t] = I8(z < 0? 0 : -sp[z]);
			}
			for (t = 2; t < tlen; ++t) { /* acceptor: what ends at position t */
				int z = 3;
				const uint8_t c0 = target[t-2], c1 = target[t-1], c2 = target[t];
				if (flag & MM2O_EZ_SPLICE_FOR) {
					if (c1 == 0 && c2 == 2) z = c0 == 1 || c0 == 3? -1 : 0;  /* [CT]AG */
					else if (c1 == 0 && c2 == 1) z = 2;                       /* AC */
				} else if (flag & MM2O_EZ_SPLICE_REV) {
					if (c1 == 0 && c2 == 1) z = c0 == 1 || c0 == 3? -1 : 0;  /* AC: reverse complement of GT */
					else if (c1 == 2 && c2 == 1) z = 1;                       /* GC */
					else if (c1 == 0 && c2 == 3) z = 2;                       /* AT */
				}
				acceptor[t] = I8(z < 0? 0 : -sp[z]);
			}
		} else { /* the sequences were reversed by the caller (left extension): signals are read backwards */
			for (t = 0; t < tlen - 4; ++t) {
				int z = 3;
				const uint8_t c1 = target[t+1], c2 = target[t+2], c3 = target[t+3];
				if (flag & MM2O_EZ_SPLICE_FOR) {
					if (c1 == 2 && c2 == 0) z = c3 == 1 || c3 == 3? -1 : 0;  /* GA = reversed AG */
					else if (c1 == 1 && c2 == 0) z = 2;                       /* CA = reversed AC */
				} else if (flag & MM2O_EZ_SPLICE_REV) {
					if (c1 == 1 && c2 == 0) z = c3 == 1 || c3 == 3? -1 : 0;  /* CA = complement of GT */
					else if (c1 == 1 && c2 == 2) z = 1;                       /* CG */
					else if (c1 == 3 && c2 == 0) z = 2;                       /* TA */
				}
				donor[t] = I8(z < 0? 0 : -sp[z]);
			}
			for (t = 2; t < tlen; ++t) {
				int z = 3;
				const uint8_t c0 = target[t-2], c1 = target[t-1], c2 = target[t];
				if (flag & MM2O_EZ_SPLICE_FOR) {
					if (c1 == 3 && c2 == 2) z = c0 == 0 || c0 == 2? -1 : 0;  /* TG = reversed GT */
					else if (c1 == 1 && c2 == 2) z = 1;                       /* CG */
					else if (c1 == 3 && c2 == 0) z = 2;                       /* TA */
				} else if (flag & MM2O_EZ_SPLICE_REV) {
					if (c1 == 3 && c2 == 1) z = c0 == 0 || c0 == 2? -1 : 0;  /* TC = complement of AG */
					else if (c1 == 3 && c2 == 2) z = 2;                       /* TG */
				}
				acceptor[t] = I8(z < 0? 0 : -sp[z]);
			}
		}
	}
	if (junc && (flag & MM2O_EZ_SPLICE_SCORE)) { /* junc[] carries per-site scores (:213-219) */
		uint8_t donor_val = !!(flag & MM2O_EZ_SPLICE_FOR) == !(flag & MM2O_EZ_REV_CIGAR)? 0 : 1;
		for (t = 0; t < tlen - 1; ++t)
			donor[t] = I8(donor[t] + (junc[t+1] == 0xff || (junc[t+1] & 1) != donor_val? -junc_pen : (int8_t)(junc[t+1] >> 1) - (int8_t)64));
		for (t = 0; t < tlen - 1; ++t)
			acceptor[t] = I8(acceptor[t] + (junc[t+1] == 0xff || (junc[t+1] & 1) != !donor_val? -junc_pen : (int8_t)(junc[t+1] >> 1) - (int8_t)64));
	} else if (junc) { /* junc[] marks annotated splice sites (:220-241) */
		const int fo = !!(flag & MM2O_EZ_SPLICE_FOR), re = !!(flag & MM2O_EZ_SPLICE_REV);
		if (!(flag & MM2O_EZ_REV_CIGAR)) {
			for (t = 0; t < tlen - 1; ++t) if ((fo && (junc[t+1] & 1)) || (re && (junc[t+1] & 8))) donor[t] = I8(donor[t] + junc_bonus);
			for (t = 0; t < tlen; ++t)     if ((fo && (junc[t] & 2))   || (re && (junc[t] & 4)))   acceptor[t] = I8(acceptor[t] + junc_bonus);
		} else {
			for (t = 0; t < tlen - 1; ++t) if ((fo && (junc[t+1] & 2)) || (re && (junc[t+1] & 4))) donor[t] = I8(donor[t] + junc_bonus);
			for (t = 0; t < tlen; ++t)     if ((fo && (junc[t] & 1))   || (re && (junc[t] & 8)))   acceptor[t] = I8(acceptor[t] + junc_bonus);
		}
	}
}

void mm2o_exts2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
				int q_, int e_, int q2_, int noncan_, int zdrop, int end_bonus, int junc_bonus_, int junc_pen_, int flag,
				const uint8_t *junc, mm2o_ez_t *ez)
{
	const int8_t q = I8(q_), e = I8(e_), q2 = I8(q2_), noncan = I8(noncan_), junc_bonus = I8(junc_bonus_), junc_pen = I8(junc_pen_);
	int r, t, qe = q + e, n_col, tlen16, qlen16, last_st, last_en, min_sc, long_thres, long_diff;
	const int with_cigar = !(flag & MM2O_EZ_SCORE_ONLY), approx_max = !!(flag & MM2O_EZ_APPROX_MAX), right = !!(flag & MM2O_EZ_RIGHT);
	int8_t *u, *v, *x, *y, *x2, *donor, *acceptor, *s, *ou, *ov, *ox, *oy, *ox2;
	int8_t sc_mch, sc_mis, sc_N;
	uint8_t *sf, *qr, *p = 0;
	int *off = 0, *off_end = 0;
	int32_t *H = 0, H0 = 0, last_H0_t = 0;

	ez->cigar = 0, ez->m_cigar = 0;
	s_reset(ez);
	if (m <= 1 || qlen <= 0 || tlen <= 0 || q2 <= q + e) return; /* :71 */
	sc_mch = mat[0], sc_mis = mat[1];
	sc_N = mat[m*m-1] == 0? I8(-e) : mat[m*m-1]; /* :80 */
	tlen16 = (tlen + 15) / 16 * 16, qlen16 = (qlen + 15) / 16 * 16;
	n_col = (((qlen < tlen? qlen : tlen) + 15) / 16 + 1) * 16; /* :84, bytes */
	for (t = 1, min_sc = mat[1]; t < m * m; ++t) min_sc = min_sc < mat[t]? min_sc : mat[t];
	if (-min_sc > 2 * (q + e)) return; /* :90 */
	long_thres = (q2 - q) / e - 1; /* :92-95 */
	if (q2 > q + e + long_thres * e) ++long_thres;
	long_diff = long_thres * e - (q2 - q);

	u = (int8_t*)calloc((size_t)tlen16 * 13, 1);
	v = u + tlen16, x = v + tlen16, y = x + tlen16, x2 = y + tlen16, donor = x2 + tlen16, acceptor = donor + tlen16, s = acceptor + tlen16;
	ou = s + tlen16, ov = ou + tlen16, ox = ov + tlen16, oy = ox + tlen16, ox2 = oy + tlen16;
	memset(u, -q - e, (size_t)tlen16 * 4); /* u, v, x, y (:102) */
	memset(x2, -q2, tlen16);
	sf = (uint8_t*)calloc(tlen16 + qlen16 + 32, 1);
	qr = sf + tlen16;
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);
	if (!approx_max) {
		H = (int32_t*)malloc((size_t)tlen16 * 4);
		for (t = 0; t < tlen16; ++t) H[t] = MM2O_KSW_NEG_INF;
	}
	if (with_cigar) {
		p = (uint8_t*)malloc((size_t)(qlen + tlen - 1) * n_col + 16);
		off = (int*)malloc((size_t)(qlen + tlen - 1) * sizeof(int) * 2);
		off_end = off + qlen + tlen - 1;
	}
	splice_signals(tlen, target, flag, noncan, junc_bonus, junc_pen, junc, donor, acceptor, tlen16);

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, x21, v1;
		const uint8_t *qrr = qr + (qlen - 1 - r);
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		if (st > 0) { /* :243-249 */
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = I8(-q - e), x21 = I8(-q2), v1 = I8(-q - e);
		} else {
			x1 = I8(-q - e), x21 = I8(-q2);
			v1 = r == 0? I8(-q - e) : r < long_thres? I8(-e) : r == long_thres? I8(long_diff) : 0;
		}
		if (en >= r) {
			y[r] = I8(-q - e);
			u[r] = r == 0? I8(-q - e) : r < long_thres? I8(-e) : r == long_thres? I8(long_diff) : 0;
		}
		if (!(flag & MM2O_EZ_GENERIC_SC)) { /* :255-275: 16-byte chunks from st0 */
			int t0;
			for (t0 = st0; t0 <= en0; t0 += 16)
				for (t = t0; t < t0 + 16 && t < tlen16; ++t) {
					uint8_t sq = sf[t], sr = qrr[t];
					int8_t sc = sq == sr? sc_mch : sc_mis;
					if (sq == (uint8_t)(m - 1) || sr == (uint8_t)(m - 1)) sc = sc_N;
					s[t] = sc;
				}
		} else for (t = st0; t <= en0; ++t) s[t] = mat[sf[t] * m + qrr[t]];
		memcpy(ou, u, tlen16), memcpy(ov, v, tlen16), memcpy(ox, x, tlen16), memcpy(oy, y, tlen16), memcpy(ox2, x2, tlen16);
		if (with_cigar) off[r] = st, off_end[r] = en;
		for (t = st; t <= en; ++t) { /* :36-66 and the three variants :283-380 */
			int8_t z = s[t], xt1, vt1, x2t1, ut, a, b, a2, a2a, tmp, d = 0, dn;
			xt1 = t == st? x1 : ox[t - 1];
			vt1 = t == st? v1 : ov[t - 1];
			x2t1 = t == st? x21 : ox2[t - 1];
			ut = ou[t];
			a = I8(xt1 + vt1), b = I8(oy[t] + ut), a2 = I8(x2t1 + vt1), a2a = I8(a2 + acceptor[t]);
			if (!with_cigar || !right) {
				d = a > z? 1 : 0;   z = z > a? z : a;
				d = b > z? 2 : d;   z = z > b? z : b;
				d = a2a > z? 3 : d; z = z > a2a? z : a2a;
			} else {
				d = z > a? 0 : 1;   z = z > a? z : a;
				d = z > b? d : 2;   z = z > b? z : b;
				d = z > a2a? d : 3; z = z > a2a? z : a2a;
			}
			u[t] = I8(z - vt1), v[t] = I8(z - ut);
			tmp = I8(z - q); a = I8(a - tmp), b = I8(b - tmp);
			a2 = I8(a2 - I8(z - q2));
			dn = donor[t];
			if (!with_cigar || !right) {
				x[t] = I8((a > 0? a : 0) - qe); if (a > 0) d |= 0x08;
				y[t] = I8((b > 0? b : 0) - qe); if (b > 0) d |= 0x10;
				x2[t] = I8((a2 > dn? a2 : dn) - q2); if (a2 > dn) d |= 0x20;
			} else {
				x[t] = I8((0 > a? 0 : a) - qe); if (!(0 > a)) d |= 0x08;
				y[t] = I8((0 > b? 0 : b) - qe); if (!(0 > b)) d |= 0x10;
				x2[t] = I8((dn > a2? dn : a2) - q2); if (!(dn > a2)) d |= 0x20;
			}
			if (with_cigar) p[(size_t)r * n_col + (t - st)] = (uint8_t)d;
		}
		if (!approx_max) { /* :381-430, the 4-lane arg-max as in mm2o_extd2 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				for (t = st0; t < en1; t += 4)
					for (i = 0; i < 4; ++i) {
						H[t + i] += v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				for (i = 0; i < 4; ++i) if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) {
					H[t] += (int32_t)v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en0;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (s_zdrop(ez, max_H, r, max_t, zdrop, 0)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* :431-447 */
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += v[last_H0_t];
				else ++last_H0_t, H0 += u[last_H0_t];
			} else H0 = v[0] - qe, last_H0_t = 0;
			if ((flag & MM2O_EZ_APPROX_DROP) && s_zdrop(ez, H0, r, last_H0_t, zdrop, 0)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(u); free(sf); free(H);
	if (with_cigar) { /* :452-463: introns become N once they are at least long_thres long */
		const int rev_cigar = !!(flag & MM2O_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & MM2O_EZ_EXTZ_ONLY)) s_backtrack(ez, rev_cigar, long_thres, p, off, off_end, n_col, tlen - 1, qlen - 1);
		else if (!ez->zdropped && (flag & MM2O_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) {
			ez->reach_end = 1;
			s_backtrack(ez, rev_cigar, long_thres, p, off, off_end, n_col, ez->mqe_t, qlen - 1);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) s_backtrack(ez, rev_cigar, long_thres, p, off, off_end, n_col, ez->max_t, ez->max_q);
		free(p); free(off);
	}
}
