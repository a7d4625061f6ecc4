// minimap2_b200/csrc/ksw_fast.cu
//
// K3 fast path: global dual-affine alignment with traceback for the gap-fill shape that dominates map-ont
// (align.c:803-872 first pass: KSW_EZ_APPROX_MAX, w >= max(qlen,tlen) so the band never clips, ~230 x 230 cells,
// ~41 such calls per 10 kb read = 95% of all DP cells).
//
// Same recurrence and tie rules as ksw_extd2_sse (ksw2_extd2_sse.c:39-66, 232-273), but the per-cell difference values
// are pure functions of the cell's two predecessors, so any evaluation order gives the reference's bytes as long as the
// band is not clipping (then no stale SSE lane is ever consumed). That freedom is used for a register-resident
// wavefront: lane l of a warp owns C consecutive target columns, all lanes walk the query rows with a skew of one row
// per lane (lane l is at row s-l in step s), the vertical state (u,y,y2 per column) never leaves registers, the
// horizontal state (v,x,x2) is handed to the next lane with one packed shuffle per step, and the only memory traffic is
// the 1 B/cell traceback row (one coalesced 8/16-byte store per lane per step). 88% of lane-steps do useful cells
// (qlen/(qlen+31)) versus ~60% for an anti-diagonal sweep, and there is no shared-memory round trip per cell.
// Valid cells never leave the int8 range (guaranteed by mm_check_opt: (q+e)+(q2+e2) <= 127), so int32 arithmetic
// reproduces the reference's wrapping int8 arithmetic bit for bit.
//
// Jobs that do not qualify (exact max / extension / right-aligned / clipped band / long) go to the universal kernel
// in ksw_extd2.cu, which emulates the SSE lane semantics exactly.
#include "mmb_internal.h"
#include <algorithm>

#define KSW_NEG_INF (-0x40000000)

namespace {

struct FastArgs {
	const mmb_ksw_job_t *jobs;
	const int *order;
	int n;
	const uint8_t *query;
	const void *target;
	int t_packed;
	mmb_ksw_res_t *res;
	uint32_t *cigar;
	long long cigar_cap;
	unsigned long long *cigar_used;
	uint8_t *pws; size_t pws_stride;      // traceback rows: qlen x (32*C) bytes per worker
	uint32_t *cigws; size_t cigws_stride;
	int *counter;
	int qmax;                             // shared-memory query staging capacity per warp
	int8_t mch, mis, scn;
	int8_t q, e, q2, e2;
	int long_thres, long_diff;
	int8_t mat[25];
};

__device__ __forceinline__ uint8_t fetch_t(const void *target, int packed, long long idx)
{
	if (packed) {
		const uint32_t *S = (const uint32_t*)target;
		return (uint8_t)(S[idx >> 3] >> ((idx & 7) << 2) & 0xf);
	}
	return ((const uint8_t*)target)[idx];
}

__device__ __forceinline__ void push_cig(uint32_t *cig, int &n, uint32_t op, int len)
{
	if (n == 0 || op != (cig[n - 1] & 0xf)) cig[n++] = (uint32_t)len << 4 | op;
	else cig[n - 1] += (uint32_t)len << 4;
}

template<int C>
__global__ void __launch_bounds__(128) ksw_fast_kernel(FastArgs A)
{
	extern __shared__ uint8_t smem[];
	const int lane = threadIdx.x & 31, wk = threadIdx.x >> 5, nwk = blockDim.x >> 5;
	const int worker = blockIdx.x * nwk + wk;
	uint8_t *qs = smem + (size_t)wk * (A.qmax + 1024);   // staged query bytes
	uint8_t *tile = qs + A.qmax;                          // 32x32 traceback window
	uint8_t *p = A.pws + A.pws_stride * worker;
	uint32_t *cig = A.cigws + A.cigws_stride * worker;
	constexpr int CP = C <= 4? 4 : C <= 8? 8 : 16;        // bytes reserved per lane per row (one aligned vector store)
	const int W = 32 * CP;                                // traceback row stride
	const int q = A.q, e = A.e, q2 = A.q2, e2 = A.e2, qe = q + e, qe2 = q2 + e2;
	const int mch = A.mch, mis = A.mis, scn = A.scn;
	const unsigned full = 0xffffffffu;

	#pragma unroll 1
	for (;;) {
		int slot = 0;
		if (lane == 0) slot = atomicAdd(A.counter, 1);
		slot = __shfl_sync(full, slot, 0);
		if (slot >= A.n) break;
		const int jid = A.order[slot];
		const mmb_ksw_job_t jb = A.jobs[jid];
		const int qlen = jb.qlen, tlen = jb.tlen;
		const bool generic = (jb.flag & MMB_KSW_GENERIC_SC) != 0;
		// stage the query (row j -> base) and this lane's target columns
		for (int i = lane; i < qlen; i += 32) {
			uint8_t c = A.query[jb.q_start + (long long)i * jb.q_step];
			if ((jb.flag & MMB_JOB_Q_COMP) && c < 4) c = 3 - c;
			qs[i] = c;
		}
		const int t0 = lane * C;
		int tb[C], u[C], y[C], y2[C], msc[C];
		#pragma unroll
		for (int c = 0; c < C; ++c) {
			const int t = t0 + c;
			tb[c] = t < tlen? fetch_t(A.target, A.t_packed, jb.t_start + (long long)t * jb.t_step) : 4;
			// top boundary (ksw2_extd2_sse.c:159-162): the cell above row 0 in column t carries the gap-open ramp of diagonal r=t
			u[c] = t == 0? -q - e : t < A.long_thres? -e : t == A.long_thres? A.long_diff : -e2;
			y[c] = -q - e, y2[c] = -q2 - e2;
			msc[c] = tb[c] == 4? scn : mis; // mismatch score of this column (ambiguous target base scores sc_N)
		}
		__syncwarp();
		int hcol0 = 0, hrow = 0;           // score pieces: lane 0 sums v down column 0, every lane sums u along the last row
		int pk_prev = 0;                   // (v,x,x2) of this lane's last column, packed, from the previous step
		const int n_steps = qlen + 31;
		#pragma unroll 1
		for (int s = 0; s < n_steps; ++s) {
			const int j = s - lane;
			// hand the right edge of lane l-1 (same query row) to lane l
			int pk_in = __shfl_up_sync(full, pk_prev, 1);
			if (j >= 0 && j < qlen && t0 < tlen) {
				int vl, xl, x2l;
				if (lane == 0) { // left boundary (ksw2_extd2_sse.c:149-158 with st == 0): diagonal r = j
					xl = -q - e, x2l = -q2 - e2;
					vl = j == 0? -q - e : j < A.long_thres? -e : j == A.long_thres? A.long_diff : -e2;
				} else {
					vl = (int)(int8_t)(pk_in & 0xff), xl = (int)(int8_t)(pk_in >> 8 & 0xff), x2l = (int)(int8_t)(pk_in >> 16 & 0xff);
				}
				const int qb = qs[j];
				int hc0 = 0;
				uint32_t dw[(C + 3) / 4];
				#pragma unroll
				for (int k = 0; k < (C + 3) / 4; ++k) dw[k] = 0;
				// Lean cell update (~40 integer ops): the 4 "open/extend" maxima are folded with the -(q+e) shift
				// (x' = max(a - (z-q) - qe, -qe)), rows with an ambiguous query base take the generic-score path.
				const int nqe = -qe, nqe2 = -qe2;
				if (!generic && qb < 4) {
					#pragma unroll
					for (int c = 0; c < C; ++c) {
						const int sc = tb[c] == qb? mch : msc[c];
						const int uc = u[c];
						int a = xl + vl, b = y[c] + uc, a2 = x2l + vl, b2 = y2[c] + uc;
						int z = sc, d = 0;
						if (a > z) d = 1;  z = max(z, a);      // ksw2_extd2_sse.c:235-243: strict '>' keeps the earlier state
						if (b > z) d = 2;  z = max(z, b);
						if (a2 > z) d = 3; z = max(z, a2);
						if (b2 > z) d = 4; z = max(z, b2);
						z = min(z, mch);
						const int un = z - vl, vn = z - uc;
						const int t1 = z - q + qe, t2 = z - q2 + qe2;
						a -= t1, b -= t1, a2 -= t2, b2 -= t2;           // = (a - (z-q)) - qe : value before the max(.,0) shifted by -qe
						if (a > nqe) d |= 0x08;
						if (b > nqe) d |= 0x10;
						if (a2 > nqe2) d |= 0x20;
						if (b2 > nqe2) d |= 0x40;
						xl = max(a, nqe), y[c] = max(b, nqe), x2l = max(a2, nqe2), y2[c] = max(b2, nqe2);
						u[c] = un, vl = vn;
						dw[c >> 2] |= (uint32_t)d << ((c & 3) * 8);
						if (c == 0) hc0 = vn;
					}
				} else {
					#pragma unroll
					for (int c = 0; c < C; ++c) {
						int sc;
						if (!generic) sc = (tb[c] == 4 || qb == 4)? scn : tb[c] == qb? mch : mis;
						else sc = A.mat[tb[c] * 5 + qb];
						const int uc = u[c];
						int a = xl + vl, b = y[c] + uc, a2 = x2l + vl, b2 = y2[c] + uc;
						int z = sc, d = 0;
						if (a > z) d = 1;  z = max(z, a);
						if (b > z) d = 2;  z = max(z, b);
						if (a2 > z) d = 3; z = max(z, a2);
						if (b2 > z) d = 4; z = max(z, b2);
						z = min(z, mch);
						const int un = z - vl, vn = z - uc;
						const int t1 = z - q + qe, t2 = z - q2 + qe2;
						a -= t1, b -= t1, a2 -= t2, b2 -= t2;
						if (a > nqe) d |= 0x08;
						if (b > nqe) d |= 0x10;
						if (a2 > nqe2) d |= 0x20;
						if (b2 > nqe2) d |= 0x40;
						xl = max(a, nqe), y[c] = max(b, nqe), x2l = max(a2, nqe2), y2[c] = max(b2, nqe2);
						u[c] = un, vl = vn;
						dw[c >> 2] |= (uint32_t)d << ((c & 3) * 8);
						if (c == 0) hc0 = vn;
					}
				}
				if (j == qlen - 1) { // H(t,qlen-1) = H(t-1,qlen-1) + u(t,qlen-1): sum u along the last row
					#pragma unroll
					for (int c = 0; c < C; ++c) if (t0 + c >= 1 && t0 + c < tlen) hrow += u[c];
				}
				if (lane == 0) hcol0 += j == 0? hc0 - qe : hc0; // H(0,0) = v - (q+e) (:366); then H(0,j) = H(0,j-1) + v(0,j)
				pk_prev = (vl & 0xff) | (xl & 0xff) << 8 | (x2l & 0xff) << 16;
				uint8_t *prow = p + (size_t)j * W + lane * CP;
				if (CP == 4) *(uint32_t*)prow = dw[0];
				else if (CP == 8) *(uint2*)prow = make_uint2(dw[0], dw[1 % ((C + 3) / 4)]);
				else *(uint4*)prow = make_uint4(dw[0], dw[1 % ((C + 3) / 4)], dw[2 % ((C + 3) / 4)], dw[3 % ((C + 3) / 4)]);
			}
		}
		// score = H(tlen-1, qlen-1): down column 0, then along the last row
		int score = hrow;
		for (int o = 16; o > 0; o >>= 1) score += __shfl_xor_sync(full, score, o);
		score += __shfl_sync(full, hcol0, 0);
		__syncwarp();
		// ---- traceback from (tlen-1, qlen-1) (ksw2.h:130-162; no forced states: the band is never clipped) ----
		int n = 0, i = tlen - 1, jj = qlen - 1, state = 0;
		for (;;) {
			int go = i >= 0 && jj >= 0;
			go = __shfl_sync(full, go, 0);
			if (!go) break;
			const int jt = __shfl_sync(full, jj, 0), it = __shfl_sync(full, i, 0);
			{ // lane k stages row jt-k, columns it-31..it
				const int rr = jt - lane;
				if (rr >= 0) {
					const uint8_t *prow = p + (size_t)rr * W;
					#pragma unroll 8
					for (int c = 0; c < 32; ++c) {
						const int col = it - 31 + c;
						tile[lane * 32 + c] = col >= 0? prow[(col / C) * CP + col % C] : 0;
					}
				}
			}
			__syncwarp();
			if (lane == 0) {
				while (i >= 0 && jj >= 0 && jt - jj < 32 && it - i < 32) {
					const uint32_t tmp = tile[(jt - jj) * 32 + (i - (it - 31))];
					if (state == 0) state = tmp & 7;
					else if (!(tmp >> (state + 2) & 1)) state = 0;
					if (state == 0) state = tmp & 7;
					if (state == 0) push_cig(cig, n, 0, 1), --i, --jj;
					else if (state == 1 || state == 3) push_cig(cig, n, 2, 1), --i;
					else push_cig(cig, n, 1, 1), --jj;
				}
			}
			__syncwarp();
		}
		if (lane == 0) {
			if (i >= 0) push_cig(cig, n, 2, i + 1);
			if (jj >= 0) push_cig(cig, n, 1, jj + 1);
		}
		n = __shfl_sync(full, n, 0);
		unsigned long long coff = 0;
		if (n > 0) {
			if (lane == 0) coff = atomicAdd(A.cigar_used, (unsigned long long)n);
			coff = __shfl_sync(full, coff, 0);
			const bool rev_cigar = (jb.flag & MMB_KSW_REV_CIGAR) != 0;
			if ((long long)(coff + n) <= A.cigar_cap)
				for (int k = lane; k < n; k += 32) A.cigar[coff + k] = rev_cigar? cig[k] : cig[n - 1 - k];
		}
		if (lane == 0) {
			mmb_ksw_res_t o;
			o.max = 0, o.zdropped = 0, o.max_q = o.max_t = o.mqe_t = o.mte_q = -1, o.mqe = o.mte = KSW_NEG_INF;
			o.score = score, o.n_cigar = n, o.reach_end = 0, o.cigar_off = (uint32_t)coff;
			A.res[jid] = o;
		}
		__syncwarp();
	}
}

} // namespace

bool mmb_ksw_fast_eligible(const mmb_ksw_job_t &j)
{
	const int f = j.flag;
	if (f & (MMB_JOB_LL | MMB_KSW_SCORE_ONLY | MMB_KSW_RIGHT | MMB_KSW_APPROX_DROP | MMB_KSW_EXTZ_ONLY)) return false;
	if (!(f & MMB_KSW_APPROX_MAX)) return false;
	if (j.qlen <= 0 || j.tlen <= 0 || j.tlen > 512 || j.qlen > 2048) return false;
	const int w = j.w < 0? std::max(j.qlen, j.tlen) : j.w;
	return w >= std::max(j.qlen, j.tlen);
}

// Launches the fast kernel over the eligible jobs listed in `idx` (indices into the batch). Asynchronous on ctx->stream
// except for the queue-order upload, which is synchronised before returning.
void mmb_ksw_fast_launch(mmb_ctx_t *ctx, const mmb_ksw_score_t *sc, const std::vector<int> &idx, const mmb_ksw_job_t *h_jobs, const mmb_ksw_job_t *d_jobs,
						 const uint8_t *d_query, const void *d_target, int t_packed, mmb_ksw_res_t *d_res, uint32_t *d_cigar, int64_t cigar_cap,
						 unsigned long long *d_cigar_used, DevBuf &pws, DevBuf &cigws, DevBuf &orderbuf)
{
	if (idx.empty()) return;
	FastArgs A;
	A.jobs = d_jobs, A.query = d_query, A.target = d_target, A.t_packed = t_packed, A.res = d_res;
	A.cigar = d_cigar, A.cigar_cap = cigar_cap, A.cigar_used = d_cigar_used;
	int8_t q = sc->q, e = sc->e, q2 = sc->q2, e2 = sc->e2;
	if (q2 + e2 < q + e) std::swap(q, q2), std::swap(e, e2);
	A.q = q, A.e = e, A.q2 = q2, A.e2 = e2;
	A.mch = sc->mat[0], A.mis = sc->mat[1], A.scn = sc->mat[24] == 0? (int8_t)(-e2) : sc->mat[24];
	for (int i = 0; i < 25; ++i) A.mat[i] = sc->mat[i];
	int lt = e != e2? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + lt * e2 > q + e + lt * e) ++lt;
	A.long_thres = lt, A.long_diff = lt * (e - e2) - (q2 - q) - e2;
	// column strips: the smallest C with 32*C >= tlen keeps the idle-lane fraction low (C is a template parameter)
	static const int CW[] = { 2, 4, 6, 7, 8, 9, 10, 12, 14, 16 };
	static const int CPW[] = { 4, 4, 8, 8, 8, 16, 16, 16, 16, 16 };
	const int NC = sizeof(CW) / sizeof(CW[0]);
	std::vector<int> v[NC];
	for (int i : idx) {
		int k = 0;
		while (k < NC - 1 && 32 * CW[k] < h_jobs[i].tlen) ++k;
		v[k].push_back(i);
	}
	int *d_order_all = (int*)orderbuf.reserve((idx.size() + NC + 1) * sizeof(int));
	size_t order_off = 0;
	for (int k = 0; k < NC; ++k) {
		if (v[k].empty()) continue;
		std::sort(v[k].begin(), v[k].end(), [&](int a, int b) {
			int64_t ca = (int64_t)h_jobs[a].qlen * h_jobs[a].tlen, cb = (int64_t)h_jobs[b].qlen * h_jobs[b].tlen;
			return ca != cb? ca > cb : a < b; });
		int maxq = 1, maxsum = 2;
		for (int i : v[k]) maxq = std::max(maxq, h_jobs[i].qlen), maxsum = std::max(maxsum, h_jobs[i].qlen + h_jobs[i].tlen);
		const int C = CW[k], W = 32 * CPW[k];
		A.qmax = (maxq + 15) / 16 * 16;
		const int threads = 128, nwk = threads / 32;
		const size_t smem = (size_t)(A.qmax + 1024) * nwk;
		void (*kern)(FastArgs) = nullptr;
		switch (C) {
		case 2: kern = ksw_fast_kernel<2>; break;   case 4: kern = ksw_fast_kernel<4>; break;
		case 6: kern = ksw_fast_kernel<6>; break;   case 7: kern = ksw_fast_kernel<7>; break;
		case 8: kern = ksw_fast_kernel<8>; break;   case 9: kern = ksw_fast_kernel<9>; break;
		case 10: kern = ksw_fast_kernel<10>; break; case 12: kern = ksw_fast_kernel<12>; break;
		case 14: kern = ksw_fast_kernel<14>; break; default: kern = ksw_fast_kernel<16>; break;
		}
		int cta_per_sm = 1;
		MMB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cta_per_sm, kern, threads, smem));
		if (cta_per_sm < 1) { fprintf(stderr, "[ERROR] ksw_fast: no occupancy\n"); abort(); }
		int grid = ctx->n_sm * cta_per_sm;
		grid = std::max(1, std::min(grid, ((int)v[k].size() + nwk - 1) / nwk));
		A.pws_stride = ((size_t)maxq * W + 255) & ~(size_t)255;
		A.cigws_stride = (size_t)maxsum + 8;
		A.pws = (uint8_t*)pws.reserve(A.pws_stride * (size_t)grid * nwk);   // may grow (cudaFree synchronises) -- rare after warm-up
		A.cigws = (uint32_t*)cigws.reserve(A.cigws_stride * 4 * (size_t)grid * nwk);
		int *d_order = d_order_all + order_off; order_off += v[k].size() + 1;
		MMB_CUDA_CHECK(cudaMemcpyAsync(d_order + 1, v[k].data(), v[k].size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
		MMB_CUDA_CHECK(cudaMemsetAsync(d_order, 0, sizeof(int), ctx->stream));
		A.counter = d_order, A.order = d_order + 1, A.n = (int)v[k].size();
		kern<<<grid, threads, smem, ctx->stream>>>(A);
		MMB_CUDA_CHECK(cudaGetLastError());
		++ctx->n_launch;
	}
}
