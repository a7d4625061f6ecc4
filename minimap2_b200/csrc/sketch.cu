// minimap2_b200/csrc/sketch.cu
//
// K1: symmetric (w,k)-minimizer sketch on sm_100a. Replaces mm_sketch (reference sketch.c:77-143).
//
// mm_sketch is a sequential state machine (ring of the last w slot values + current minimum + run length l).
// It is chunk-parallelised here: one thread owns the positions [cs, ce) of one sequence and replays the reference
// state machine from a warm-up point R0 = cs - WU with a cleared state. After the warm-up the state equals the
// sequential one provided (a) no ambiguous base lies in [R0, cs) and (b) at most 16 symmetric-k-mer skips occurred,
// because then the ring holds the true last w slot writes, the k-mer registers hold the true last k bases and the run
// length is saturated (>= w+k) in both runs, which is all the emission rules (sketch.c:117-139) look at. A chunk
// whose warm-up fails the check (N nearby, long ATAT.. stretch) or whose sequence uses HPC is replayed from the
// sequence start instead -- slower, still exact. Emissions are produced in position order per chunk, so a count pass,
// an exclusive scan and a write pass give exactly the reference's output order with no sort.
#include "mmb_internal.h"
#include "mm_algo.cuh"
#include "scan.cuh"

namespace {

struct SeqSrc {          // how bases are fetched: nt4/ASCII bytes, or the 4-bit packed index sequence
	const uint8_t *bytes;
	const uint32_t *packed;
	__device__ __forceinline__ int get(int64_t i) const {
		if (packed) return (int)mmx_seq4_get(packed, (uint64_t)i) > 3? 4 : (int)mmx_seq4_get(packed, (uint64_t)i);
		return mmx_nt4(bytes[i]);
	}
};

struct SketchArgs {
	SeqSrc src;
	const int64_t *seq_off;     // n_seq+1 base offsets of the sequences in src
	const int64_t *chunk_off;   // n_seq+1: first chunk id of each sequence
	const uint32_t *rid;        // optional per-sequence rid (NULL: rid0 + seq index)
	uint32_t rid0;
	int n_seq, w, k, is_hpc, chunk, same_rid;
	int64_t n_chunks;
	int64_t *cnt;               // per chunk count (pass 0) / exclusive offsets (pass 1 input)
	m128 *out;
};

// binary search: largest s with chunk_off[s] <= c
__device__ __forceinline__ int find_seq(const int64_t *chunk_off, int n_seq, int64_t c)
{
	int lo = 0, hi = n_seq;
	while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (chunk_off[mid] <= c) lo = mid; else hi = mid; }
	return lo;
}

#define MAXW 256

template<bool WRITE>
__global__ void __launch_bounds__(128) sketch_kernel(SketchArgs A)
{
	const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= A.n_chunks) return;
	const int s = find_seq(A.chunk_off, A.n_seq, c);
	const int64_t base = A.seq_off[s];
	const int len = (int)(A.seq_off[s + 1] - base);
	const int w = A.w, k = A.k;
	const uint32_t rid = A.rid? A.rid[s] : A.same_rid? A.rid0 : A.rid0 + (uint32_t)s;
	int cs, ce;
	if (A.is_hpc) { cs = 0, ce = len; }           // HPC: one chunk per sequence (positions jump over homopolymers)
	else { cs = (int)((c - A.chunk_off[s]) * A.chunk); ce = min(len, cs + A.chunk); }
	const uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1;
	const uint64_t MAXV = ~0ULL;
	uint64_t bx[MAXW];      // ring: x values (hash<<8|span); y is rebuilt from (pos,strand) kept in bp
	uint32_t bp[MAXW];      // ring: pos<<1|strand, 0xffffffff for "no k-mer"
	int64_t n_out = 0;
	m128 *out = WRITE? A.out + A.cnt[c] : nullptr;

	const int WU = 2 * w + 2 * k + 15;
	int start = cs - WU;
	bool warm = start > 0 && !A.is_hpc;
	if (!warm) start = 0;

	for (int attempt = 0; attempt < 2; ++attempt) {
		uint64_t kmer0 = 0, kmer1 = 0, minx = MAXV;
		uint32_t minp = 0xffffffffu;
		int l = 0, buf_pos = 0, min_pos = 0, kmer_span = 0, nskip = 0, after = 0; // after: slot writes at positions >= ce
		int tq[32], tq_front = 0, tq_cnt = 0;
		bool bad = false, checked = false;
		n_out = 0;
		for (int j = 0; j < w; ++j) bx[j] = MAXV, bp[j] = 0xffffffffu;
#define EMIT(X, P) do { int pos__ = (int)((P) >> 1); if (pos__ >= cs && pos__ < ce) { \
			if (WRITE) { out[n_out].x = (X); out[n_out].y = (uint64_t)rid << 32 | (P); } ++n_out; } } while (0)
		int i;
		for (i = start; i < len; ++i) {
			if (warm && !checked && i >= cs) { // warm-up verdict: enough non-skipped k-mers => ring, k-mer registers and run length are exact
				checked = true;
				if (nskip > 16) { bad = true; break; }
				if (l < w + k) l = w + k; // the true run length is >= w+k here; every rule compares l with thresholds <= w+k
			}
			int cch = A.src.get(base + i);
			uint64_t ix = MAXV; uint32_t ip = 0xffffffffu;
			if (cch < 4) {
				if (A.is_hpc) { // sketch.c:94-104
					int skip_len = 1;
					if (i + 1 < len && A.src.get(base + i + 1) == cch) {
						for (skip_len = 2; i + skip_len < len; ++skip_len)
							if (A.src.get(base + i + skip_len) != cch) break;
						i += skip_len - 1;
					}
					tq[(tq_cnt++ + tq_front) & 0x1f] = skip_len;
					kmer_span += skip_len;
					if (tq_cnt > k) { kmer_span -= tq[tq_front++]; tq_front &= 0x1f; --tq_cnt; }
				} else kmer_span = l + 1 < k? l + 1 : k;
				kmer0 = (kmer0 << 2 | (uint64_t)cch) & mask;
				kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)cch) << shift1;
				if (kmer0 == kmer1) { if (warm && i < cs && i >= start + k - 1) ++nskip; continue; }
				int z = kmer0 < kmer1? 0 : 1;
				++l;
				if (l >= k && kmer_span < 256) {
					ix = mmx_hash64(z? kmer1 : kmer0, mask) << 8 | (uint64_t)kmer_span;
					ip = (uint32_t)i << 1 | (uint32_t)z;
				}
			} else {
				l = 0, tq_cnt = tq_front = 0, kmer_span = 0;
				if (warm && i < cs) { bad = true; break; }
			}
			bx[buf_pos] = ix, bp[buf_pos] = ip;
			const int leff = l;
			if (leff == w + k - 1 && minx != MAXV) { // sketch.c:117-122
				for (int j = buf_pos + 1; j < w; ++j) if (minx == bx[j] && bp[j] != minp) EMIT(bx[j], bp[j]);
				for (int j = 0; j < buf_pos; ++j)     if (minx == bx[j] && bp[j] != minp) EMIT(bx[j], bp[j]);
			}
			if (ix <= minx) { // sketch.c:123-125
				if (leff >= w + k && minx != MAXV) EMIT(minx, minp);
				minx = ix, minp = ip, min_pos = buf_pos;
			} else if (buf_pos == min_pos) { // sketch.c:126-138
				if (leff >= w + k - 1 && minx != MAXV) EMIT(minx, minp);
				minx = MAXV;
				for (int j = buf_pos + 1; j < w; ++j) if (minx >= bx[j]) minx = bx[j], minp = bp[j], min_pos = j;
				for (int j = 0; j <= buf_pos; ++j)    if (minx >= bx[j]) minx = bx[j], minp = bp[j], min_pos = j;
				if (leff >= w + k - 1 && minx != MAXV) {
					for (int j = buf_pos + 1; j < w; ++j) if (minx == bx[j] && minp != bp[j]) EMIT(bx[j], bp[j]);
					for (int j = 0; j <= buf_pos; ++j)    if (minx == bx[j] && minp != bp[j]) EMIT(bx[j], bp[j]);
				}
			}
			if (++buf_pos == w) buf_pos = 0;
			if (i >= ce && ++after > w) break; // everything with pos < ce has left the window
		}
		if (!bad) {
			if (i >= len && minx != MAXV) EMIT(minx, minp); // sketch.c:141-142
			break;
		}
		warm = false, start = 0; // replay from the sequence start: exact by construction
	}
#undef EMIT
	if (!WRITE) A.cnt[c] = n_out;
}

__global__ void chunk_count_kernel(const int64_t *seq_off, int n_seq, int chunk, int is_hpc, int64_t *nchunk)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_seq) {
		int64_t len = seq_off[i + 1] - seq_off[i];
		nchunk[i] = len <= 0? 0 : is_hpc? 1 : (len + chunk - 1) / chunk;
	}
}

// per-sequence minimizer counts from per-chunk exclusive offsets
__global__ void seq_count_kernel(const int64_t *chunk_off, const int64_t *cnt_excl, int64_t total, int64_t n_chunks, int n_seq, int64_t *seq_mz_off)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i <= n_seq) {
		int64_t c = i < n_seq? chunk_off[i] : n_chunks;
		seq_mz_off[i] = c < n_chunks? cnt_excl[c] : total;
	}
}

} // namespace

// Device-level launcher. d_seq_off: n_seq+1 int64 base offsets. Results: *d_out (grown as needed) holds all
// minimizers in (sequence, position) order; d_mz_off (n_seq+1 int64, device) their per-sequence offsets.
// Returns the total count (synchronises the stream once to size the output).
int64_t mmb_sketch_device(mmb_ctx_t *ctx, const uint8_t *d_bytes, const uint32_t *d_packed, const int64_t *d_seq_off, int n_seq,
						  const uint32_t *d_rid, uint32_t rid0, int w, int k, int is_hpc, int64_t total_bases,
						  DevBuf &out, int64_t *d_mz_off, DevBuf &tmp_a, DevBuf &tmp_b, int same_rid)
{
	if (n_seq <= 0) return 0;
	if (!(w > 0 && w < 256 && k > 0 && k <= 28)) { fprintf(stderr, "[ERROR] mm_sketch: invalid w=%d k=%d\n", w, k); abort(); }
	ProfScope prof(ctx, MMB_PROF_SKETCH, (uint64_t)total_bases);
	const int WU = 2 * w + 2 * k + 15;
	int chunk = 512;
	while (chunk < 6 * WU) chunk <<= 1;
	// chunk table
	int64_t *d_chunk_off = tmp_a.as<int64_t>((size_t)n_seq + 1);
	chunk_count_kernel<<<(n_seq + 255) / 256, 256, 0, ctx->stream>>>(d_seq_off, n_seq, chunk, is_hpc, d_chunk_off);
	++ctx->n_launch;
	int64_t n_chunks = mmb_exclusive_scan_i64(ctx, d_chunk_off, n_seq, true);
	if (n_chunks == 0) { MMB_CUDA_CHECK(cudaMemsetAsync(d_mz_off, 0, sizeof(int64_t) * (n_seq + 1), ctx->stream)); return 0; }
	int64_t *d_cnt = tmp_b.as<int64_t>((size_t)n_chunks + 1);
	SketchArgs A;
	A.src.bytes = d_bytes, A.src.packed = d_packed;
	A.seq_off = d_seq_off, A.chunk_off = d_chunk_off, A.rid = d_rid, A.rid0 = rid0;
	A.n_seq = n_seq, A.w = w, A.k = k, A.is_hpc = is_hpc, A.chunk = chunk, A.n_chunks = n_chunks, A.cnt = d_cnt, A.out = nullptr, A.same_rid = same_rid;
	const int threads = 128;
	const unsigned grid = (unsigned)((n_chunks + threads - 1) / threads);
	sketch_kernel<false><<<grid, threads, 0, ctx->stream>>>(A);
	MMB_CUDA_CHECK(cudaGetLastError());
	++ctx->n_launch;
	int64_t total = mmb_exclusive_scan_i64(ctx, d_cnt, n_chunks, true);
	A.out = out.as<m128>((size_t)total + 1);
	if (total > 0) {
		sketch_kernel<true><<<grid, threads, 0, ctx->stream>>>(A);
		MMB_CUDA_CHECK(cudaGetLastError());
		++ctx->n_launch;
	}
	seq_count_kernel<<<(n_seq + 1 + 255) / 256, 256, 0, ctx->stream>>>(d_chunk_off, d_cnt, total, n_chunks, n_seq, d_mz_off);
	++ctx->n_launch;
	return total;
}

extern "C" int64_t mmb_sketch_batch_host(mmb_ctx_t *ctx, int n_seq, const char *seqs, const int64_t *off, int w, int k, int is_hpc,
										 uint32_t rid0, uint64_t *out_xy, int64_t out_cap, int64_t *n_out)
{
	if (n_seq <= 0) return 0;
	MMB_CUDA_CHECK(cudaSetDevice(ctx->device));
	const int64_t tot = off[n_seq];
	uint8_t *d_seq = ctx->d_a.as<uint8_t>((size_t)tot + 16);
	int64_t *d_off = ctx->d_b.as<int64_t>((size_t)n_seq + 1);
	int64_t *d_mz_off = ctx->d_c.as<int64_t>((size_t)n_seq + 1);
	MMB_CUDA_CHECK(cudaMemcpyAsync(d_seq, seqs, tot, cudaMemcpyHostToDevice, ctx->stream));
	MMB_CUDA_CHECK(cudaMemcpyAsync(d_off, off, sizeof(int64_t) * (n_seq + 1), cudaMemcpyHostToDevice, ctx->stream));
	int64_t total = mmb_sketch_device(ctx, d_seq, nullptr, d_off, n_seq, nullptr, rid0, w, k, is_hpc, tot, ctx->d_d, d_mz_off, ctx->d_e, ctx->d_f, 0);
	if (n_out) {
		std::vector<int64_t> mo(n_seq + 1);
		MMB_CUDA_CHECK(cudaMemcpyAsync(mo.data(), d_mz_off, sizeof(int64_t) * (n_seq + 1), cudaMemcpyDeviceToHost, ctx->stream));
		MMB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
		for (int i = 0; i < n_seq; ++i) n_out[i] = mo[i + 1] - mo[i];
	}
	if (out_xy && total <= out_cap && total > 0) {
		MMB_CUDA_CHECK(cudaMemcpyAsync(out_xy, ctx->d_d.p, sizeof(m128) * total, cudaMemcpyDeviceToHost, ctx->stream));
		MMB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
	}
	return total;
}
