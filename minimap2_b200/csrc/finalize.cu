// minimap2_b200/csrc/finalize.cu -- K4: the per-hit tail of the alignment driver on the device.
//
// After the ksw2 jobs of a hit (left extension, gap fills, right extension) have run, the reference assembles the hit's CIGAR with
// mm_append_cigar (align.c:305-334), left-aligns its gaps and drops a leading I/D with mm_fix_cigar (align.c:105-181), and rescans
// every aligned base in mm_update_extra (align.c:254-303) for blen / mlen / n_ambi / dp_max. On the host that was the largest single
// cost of the replay driver (a third of its CPU time: one pass over every aligned base of every read). Here one thread per hit does
// the three steps straight from the job CIGARs that are still in the device arena, so only the final CIGAR and eight integers per
// hit cross PCIe. The running score of mm_update_extra is a sum of integers and of e * mg_log2(1+len) terms that are exact
// multiples of 2^-32 (float mantissa), so it is carried in 2^-32 fixed point -- exactly the reference's double arithmetic; a hit
// whose penalty is not representable is flagged (status 1) and redone by the host driver.
#include "pipeline.h"

namespace {

struct Acc { // the hit's two sequences as mm_update_extra sees them: query piece on its strand, target piece forward
	const uint8_t *query; const uint32_t *S;
	int64_t q0, t0; int32_t qlen, qs, rev;
	__device__ __forceinline__ int q(int x) const { // base x of the aligned query piece (align.c:1056-1061 for the reverse strand)
		const int p = qs + x;
		if (!rev) return query[q0 + p];
		const int c = query[q0 + (qlen - 1 - p)];
		return c < 4? 3 - c : 4;
	}
	__device__ __forceinline__ int t(int y) const { return (int)mmx_seq4_get(S, (uint64_t)(t0 + y)); }
};

__global__ void __launch_bounds__(128) finalize_kernel(const FinReg *regs, const FinJobRef *jobs, int n, const uint8_t *query, const uint32_t *S,
													   uint32_t *out, FinOut *res, FinPar par)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n) return;
	const FinReg R = regs[r];
	uint32_t *c = out + R.out_off;
	Acc A; A.query = query, A.S = S, A.q0 = R.q0, A.t0 = R.t0, A.qlen = R.qlen, A.qs = R.qs, A.rev = R.rev;
	FinOut o; o.status = 0, o.qshift = o.tshift = 0, o.is_spliced = 0;
	// ---- mm_append_cigar over the jobs in driver order: the first operation of a piece merges into the last one so far ----
	uint32_t nc = 0;
	for (int j = 0; j < R.n_jobs; ++j) {
		const FinJobRef J = jobs[R.job_first + j];
		if (J.n == 0) continue;
		uint32_t k = 0;
		if (nc > 0 && (c[nc - 1] & 0xf) == (J.cig[0] & 0xf)) c[nc - 1] += J.cig[0] >> 4 << 4, k = 1;
		for (; k < J.n; ++k) c[nc++] = J.cig[k];
	}
	// ---- mm_fix_cigar (align.c:105-181) ----
	int32_t qoff = 0, toff = 0;
	if (nc > 1) {
		int to_shrink = 0;
		for (uint32_t k = 0; k < nc; ++k) { // left-align gaps
			const uint32_t op = c[k] & 0xf, len = c[k] >> 4;
			if (len == 0) to_shrink = 1;
			if (op == MM_CIGAR_MATCH) toff += len, qoff += len;
			else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
				if (k > 0 && k < nc - 1 && (c[k - 1] & 0xf) == 0 && (c[k + 1] & 0xf) == 0) {
					const int prev_len = (int)(c[k - 1] >> 4);
					int l;
					if (op == MM_CIGAR_INS) { for (l = 0; l < prev_len; ++l) if (A.q(qoff - 1 - l) != A.q(qoff + len - 1 - l)) break; }
					else { for (l = 0; l < prev_len; ++l) if (A.t(toff - 1 - l) != A.t(toff + len - 1 - l)) break; }
					if (l > 0) c[k - 1] -= (uint32_t)l << 4, c[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
					if (l == prev_len) to_shrink = 1;
				}
				if (op == MM_CIGAR_INS) qoff += len; else toff += len;
			} else if (op == MM_CIGAR_N_SKIP) toff += len;
		}
		if (qoff != R.qspan || toff != R.tspan) o.status = 2; // the reference asserts this; never expected
		for (uint32_t k = 0; k + 2 < nc; ++k) { // runs like 5I6D7I become one I and one D
			if ((c[k] & 0xf) > 0 && (c[k] & 0xf) + (c[k + 1] & 0xf) == 3) {
				uint32_t l, s[3] = {0, 0, 0};
				for (l = k; l < nc; ++l) {
					const uint32_t op = c[l] & 0xf;
					if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL || c[l] >> 4 == 0) s[op < 3? op : 0] += c[l] >> 4;
					else break;
				}
				if (s[1] > 0 && s[2] > 0 && l - k > 2) {
					c[k] = s[1] << 4 | MM_CIGAR_INS;
					c[k + 1] = s[2] << 4 | MM_CIGAR_DEL;
					for (k += 2; k < l; ++k) c[k] &= 0xf;
					to_shrink = 1;
				}
				k = l;
			}
		}
		if (to_shrink) { // drop empty operations, then merge equal neighbours
			uint32_t l = 0;
			for (uint32_t k = 0; k < nc; ++k) if (c[k] >> 4 != 0) c[l++] = c[k];
			nc = l, l = 0;
			for (uint32_t k = 0; k < nc; ++k)
				if (k == nc - 1 || (c[k] & 0xf) != (c[k + 1] & 0xf)) c[l++] = c[k];
				else c[k + 1] += c[k] >> 4 << 4;
			nc = l;
		}
		if ((c[0] & 0xf) == MM_CIGAR_INS || (c[0] & 0xf) == MM_CIGAR_DEL) { // a leading I/D goes; the caller moves qs/qe or rs by the shift
			const int32_t l = (int32_t)(c[0] >> 4);
			if ((c[0] & 0xf) == MM_CIGAR_INS) o.qshift = l; else o.tshift = l;
			--nc;
			for (uint32_t k = 0; k < nc; ++k) c[k] = c[k + 1];
		}
	}
	// ---- mm_update_extra (align.c:266-297) in 2^-32 fixed point ----
	const int qsh = o.qshift, tsh = o.tshift;
	int64_t sfx = 0, maxfx = 0;
	int32_t blen = 0, mlen = 0, n_ambi_tot = 0;
	qoff = toff = 0;
	for (uint32_t k = 0; k < nc; ++k) {
		const uint32_t op = c[k] & 0xf, len = c[k] >> 4;
		if (op == MM_CIGAR_MATCH) {
			int n_ambi = 0, n_diff = 0;
			for (uint32_t l = 0; l < len; ++l) {
				const int cq = A.q(qsh + qoff + l), ct = A.t(tsh + toff + l);
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				sfx += (int64_t)par.mat[ct * 5 + cq] << 32;
				if (sfx < 0) sfx = 0;
				else maxfx = maxfx > sfx? maxfx : sfx;
			}
			blen += len - n_ambi, mlen += len - (n_ambi + n_diff), n_ambi_tot += n_ambi;
			toff += len, qoff += len;
		} else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
			int n_ambi = 0;
			if (op == MM_CIGAR_INS) { for (uint32_t l = 0; l < len; ++l) if (A.q(qsh + qoff + l) > 3) ++n_ambi; }
			else { for (uint32_t l = 0; l < len; ++l) if (A.t(tsh + toff + l) > 3) ++n_ambi; }
			blen += len - n_ambi, n_ambi_tot += n_ambi;
			const double pen = par.log_gap? par.q + (double)par.e * mmx_log2((float)(1.0 + len)) : (double)(par.q + par.e);
			const double scaled = pen * 4294967296.0;
			const int64_t pfx = (int64_t)scaled;
			if ((double)pfx != scaled || pen > 1e6 || pen < -1e6) o.status = 1;
			sfx -= pfx;
			if (sfx < 0) sfx = 0;
			if (op == MM_CIGAR_INS) qoff += len; else toff += len;
		} else if (op == MM_CIGAR_N_SKIP) o.is_spliced = 1, toff += len;
	}
	if (nc > 0 && (qoff + qsh != R.qspan || toff + tsh != R.tspan) && o.status == 0) o.status = 2;
	o.n_cigar = (int32_t)nc, o.blen = blen, o.mlen = mlen, o.n_ambi = n_ambi_tot;
	o.dp_max = (int32_t)((double)maxfx / 4294967296.0 + .499);
	res[r] = o;
}

} // namespace

void mmb_finalize_device(mmb_ctx_t *ctx, const FinReg *d_regs, const FinJobRef *d_jobs, int n_regs, const uint8_t *d_query, const uint32_t *d_S,
						 uint32_t *d_out, FinOut *d_res, const FinPar &par)
{
	if (n_regs <= 0) return;
	ProfScope ps(ctx, MMB_PROF_OTHER, (uint64_t)n_regs);
	finalize_kernel<<<(n_regs + 127) / 128, 128, 0, ctx->stream>>>(d_regs, d_jobs, n_regs, d_query, d_S, d_out, d_res, par);
	MMB_CUDA_CHECK(cudaGetLastError());
	++ctx->n_launch;
}

// Kernel-level entry with host buffers (parity tests): see include/mm_b200.h
extern "C" int mmb_tail_batch_host(mmb_ctx_t *ctx, int n_hits, const mmb_tail_hit_t *hits, int64_t n_pieces_total, const uint32_t *piece_len, const uint32_t *ops,
								   const uint8_t *query, int64_t query_len, const uint8_t *target, int64_t target_len, const int8_t *mat, int q, int e,
								   const int64_t *cig_off, mmb_tail_out_t *out, uint32_t *cigar_out)
{
	static_assert(sizeof(mmb_tail_out_t) == sizeof(FinOut), "mmb_tail_out_t mirrors FinOut");
	if (n_hits <= 0) return 0;
	MMB_CUDA_CHECK(cudaSetDevice(ctx->device));
	std::vector<int64_t> poff((size_t)n_pieces_total + 1, 0);
	for (int64_t i = 0; i < n_pieces_total; ++i) poff[i + 1] = poff[i] + piece_len[i];
	const int64_t n_ops = poff[n_pieces_total], tot_out = cig_off[n_hits];
	std::vector<uint32_t> S((size_t)(target_len + 7) / 8 + 1, 0);
	for (int64_t i = 0; i < target_len; ++i) S[i >> 3] |= (uint32_t)(target[i] & 0xf) << ((i & 7) << 2);
	uint8_t *d_q = ctx->d_a.as<uint8_t>((size_t)query_len + 16);
	uint32_t *d_S = ctx->d_b.as<uint32_t>(S.size());
	uint32_t *d_ops = ctx->d_c.as<uint32_t>((size_t)n_ops + 1);
	FinReg *d_regs = ctx->d_d.as<FinReg>((size_t)n_hits);
	FinJobRef *d_jobs = ctx->d_e.as<FinJobRef>((size_t)n_pieces_total + 1);
	uint32_t *d_out = ctx->d_f.as<uint32_t>((size_t)tot_out + 1);
	FinOut *d_res = ctx->d_g.as<FinOut>((size_t)n_hits);
	std::vector<FinReg> hr((size_t)n_hits);
	std::vector<FinJobRef> hj((size_t)n_pieces_total + 1);
	for (int64_t i = 0; i < n_pieces_total; ++i) hj[i].cig = d_ops + poff[i], hj[i].n = piece_len[i], hj[i].pad = 0;
	for (int i = 0; i < n_hits; ++i) {
		FinReg &f = hr[i];
		f.q0 = hits[i].q0, f.t0 = hits[i].t0, f.out_off = cig_off[i], f.qlen = hits[i].qlen, f.qs = hits[i].qs, f.rev = hits[i].rev;
		f.qspan = hits[i].qspan, f.tspan = hits[i].tspan, f.job_first = hits[i].piece_first, f.n_jobs = hits[i].n_pieces, f.pad = 0;
	}
	FinPar par;
	for (int i = 0; i < 25; ++i) par.mat[i] = mat[i];
	par.q = (int8_t)q, par.e = (int8_t)e, par.log_gap = 1;
	MMB_CUDA_CHECK(cudaMemcpyAsync(d_q, query, query_len, cudaMemcpyHostToDevice, ctx->stream));
	MMB_CUDA_CHECK(cudaMemcpyAsync(d_S, S.data(), S.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (n_ops) MMB_CUDA_CHECK(cudaMemcpyAsync(d_ops, ops, (size_t)n_ops * 4, cudaMemcpyHostToDevice, ctx->stream));
	MMB_CUDA_CHECK(cudaMemcpyAsync(d_regs, hr.data(), sizeof(FinReg) * (size_t)n_hits, cudaMemcpyHostToDevice, ctx->stream));
	MMB_CUDA_CHECK(cudaMemcpyAsync(d_jobs, hj.data(), sizeof(FinJobRef) * (size_t)n_pieces_total, cudaMemcpyHostToDevice, ctx->stream));
	mmb_finalize_device(ctx, d_regs, d_jobs, n_hits, d_q, d_S, d_out, d_res, par);
	MMB_CUDA_CHECK(cudaMemcpyAsync(out, d_res, sizeof(FinOut) * (size_t)n_hits, cudaMemcpyDeviceToHost, ctx->stream));
	if (tot_out) MMB_CUDA_CHECK(cudaMemcpyAsync(cigar_out, d_out, (size_t)tot_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
	MMB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
	return 0;
}
