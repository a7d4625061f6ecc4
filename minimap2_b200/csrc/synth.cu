// minimap2_b200/csrc/synth.cu -- deterministic synthetic genome / read generators on the device (bench + large-scale tests).
// Not part of the mapping path. Reproduces the SHAPES of SURVEY 8(d): uniform ACGT contigs; fixed-length reads sampled
// uniformly, 50% reverse-complemented, ONT-like errors (sub/ins/del split). Counter-based hashing makes every base a pure
// function of (seed, position), so the same data can be regenerated anywhere (and written to FASTA for the reference arm).
#include "index.h"
#include "mm_algo.cuh"
#include <cstring>
#include <cmath>
#include <string>

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) // splitmix64 finaliser
{
	x += 0x9E3779B97F4A7C15ULL;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
	return x ^ (x >> 31);
}

__global__ void genome_kernel(uint32_t *S, uint64_t n_words, uint64_t total, uint64_t seed)
{
	uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n_words) return;
	uint64_t r = mix64(seed * 0x100000001B3ULL + w), v = 0;
	for (int i = 0; i < 8; ++i) { // 8 bases per word, 2 random bits each
		uint64_t pos = w * 8 + i;
		if (pos < total) v |= ((r >> (2 * i)) & 3) << (4 * i);
	}
	S[w] = (uint32_t)v;
}

struct ReadGenArgs {
	const uint32_t *S; const uint64_t *seq_off; const uint32_t *seq_len; uint32_t n_seq; uint64_t total;
	int n_reads, read_len; uint64_t seed; float err, sub, ins;
	char *out; // n_reads * read_len ASCII
};

__global__ void reads_kernel(ReadGenArgs A)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= A.n_reads) return;
	uint64_t h = mix64(A.seed ^ (0xABCDEF12345ULL * (uint64_t)(r + 1)));
	// choose a start so that read_len*1.3 source bases fit inside one contig
	const int span = A.read_len + A.read_len / 3 + 16;
	uint64_t start = 0; uint32_t sid = 0;
	for (int tries = 0; tries < 64; ++tries) {
		h = mix64(h);
		uint64_t g = h % A.total;
		uint32_t lo = 0, hi = A.n_seq; // contig containing g
		while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (A.seq_off[mid] <= g) lo = mid; else hi = mid; }
		if (g + span <= A.seq_off[lo] + A.seq_len[lo]) { start = g, sid = lo; break; }
		start = A.seq_off[lo], sid = lo; // fallback: contig start
	}
	(void)sid;
	h = mix64(h);
	const bool rev = h & 1;
	char *o = A.out + (size_t)r * A.read_len;
	uint64_t p = start;
	int n = 0;
	const uint32_t te = (uint32_t)(A.err * 16777216.0f), ts = (uint32_t)(A.err * A.sub * 16777216.0f), ti = (uint32_t)(A.err * (A.sub + A.ins) * 16777216.0f);
	while (n < A.read_len) {
		h = mix64(h + p);
		const uint32_t u = (uint32_t)(h & 0xffffff);
		uint32_t c = mmx_seq4_get(A.S, p) & 3;
		int emit = -1;
		if (u >= te) emit = c, ++p;                                   // match
		else if (u < ts) emit = (c + 1 + ((h >> 24) % 3)) & 3, ++p;   // substitution
		else if (u < ti) emit = (h >> 24) & 3;                        // insertion (source not consumed)
		else ++p;                                                     // deletion
		if (emit >= 0) {
			const int idx = rev? A.read_len - 1 - n : n;
			const int b = rev? 3 - emit : emit;
			o[idx] = "ACGT"[b];
			++n;
		}
	}
}

// cDNA reads (SURVEY 8d "splice"): a transcript is a run of exons (100-500 bp) separated by introns (log-uniform 100 bp - 50 kb) whose
// first two / last two bases are GT / AG on the genome's forward strand (the exon is extended base by base until the genome offers the
// signal, so the genome itself is not edited); the read is the first read_len transcript bases with 3% errors (sub/ins/del 1:1:1),
// reverse-complemented for every other read (those introns then read CT..AC, the other transcript strand).
__global__ void cdna_reads_kernel(ReadGenArgs A)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= A.n_reads) return;
	uint64_t h = mix64(A.seed ^ (0x5151F00DULL * (uint64_t)(r + 1)));
	const uint64_t span = 600000; // room for the read plus its introns inside one contig
	uint64_t start = 0, end = 0;
	for (int tries = 0; tries < 64; ++tries) {
		h = mix64(h);
		uint64_t g = h % A.total;
		uint32_t lo = 0, hi = A.n_seq;
		while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (A.seq_off[mid] <= g) lo = mid; else hi = mid; }
		end = A.seq_off[lo] + A.seq_len[lo];
		if (g + span <= end) { start = g; break; }
		start = A.seq_off[lo];
	}
	h = mix64(h);
	const bool rev = h & 1;
	char *o = A.out + (size_t)r * A.read_len;
	uint64_t p = start;
	int n = 0;
	const uint32_t te = (uint32_t)(A.err * 16777216.0f), ts = te / 3, ti = 2 * (te / 3);
	while (n < A.read_len) {
		h = mix64(h + p);
		int exon_left = 100 + (int)(h % 401);
		// exon bases, extended until the genome shows GT right after the exon
		while (n < A.read_len && p + 4 < end) {
			if (exon_left <= 0 && (mmx_seq4_get(A.S, p) & 3) == 2 && (mmx_seq4_get(A.S, p + 1) & 3) == 3) break;
			--exon_left;
			h = mix64(h + p);
			const uint32_t u = (uint32_t)(h & 0xffffff);
			const uint32_t c = mmx_seq4_get(A.S, p) & 3;
			int emit = -1;
			if (u >= te) emit = c, ++p;
			else if (u < ts) emit = (c + 1 + ((h >> 24) % 3)) & 3, ++p;
			else if (u < ti) emit = (h >> 24) & 3;
			else ++p;
			if (emit >= 0) {
				const int idx = rev? A.read_len - 1 - n : n;
				o[idx] = "ACGT"[rev? 3 - emit : emit];
				++n;
			}
		}
		if (n >= A.read_len) break;
		if (p + 4 >= end) { // ran out of contig (practically unreachable with span): pad with A
			while (n < A.read_len) { o[rev? A.read_len - 1 - n : n] = 'A'; ++n; }
			break;
		}
		// intron: log-uniform length, its end moved right until the genome shows AG
		h = mix64(h + p);
		const float uu = (float)(h & 0xffffff) / 16777216.0f;
		uint64_t q = p + (uint64_t)(100.0f * expf(uu * 6.2146f)); // 100 .. 50 000
		while (q + 4 < end && !((mmx_seq4_get(A.S, q - 2) & 3) == 0 && (mmx_seq4_get(A.S, q - 1) & 3) == 2)) ++q;
		if (q + 4 >= end) q = p; // no room: no intron
		p = q;
	}
}

} // namespace

// Build an index over a synthetic genome generated on the device: n_contigs equal contigs named chr1..chrN.
extern "C" mm_idx_t *mmb_synth_index(uint64_t total_len, int n_contigs, uint64_t seed, int w, int k, int bucket_bits)
{
	mmb_ctx_t *ctx = mmb_default_ctx();
	MMB_CUDA_CHECK(cudaSetDevice(ctx->device));
	mm_idx_t *mi = (mm_idx_t*)calloc(1, sizeof(mm_idx_t));
	mi->w = w, mi->k = k, mi->b = bucket_bits, mi->flag = 0;
	mi->B = new mm_idx_bucket_s();
	mi->n_seq = n_contigs;
	mi->seq = (mm_idx_seq_t*)calloc(n_contigs + 1, sizeof(mm_idx_seq_t));
	uint64_t off = 0;
	for (int i = 0; i < n_contigs; ++i) {
		uint64_t len = total_len / n_contigs + (i < (int)(total_len % n_contigs)? 1 : 0);
		std::string nm = "chr" + std::to_string(i + 1);
		mi->seq[i].name = strdup(nm.c_str());
		mi->seq[i].offset = off, mi->seq[i].len = (uint32_t)len, mi->seq[i].is_alt = 0;
		off += len;
	}
	const uint64_t nw = (total_len + 7) / 8;
	uint32_t *d_S;
	MMB_CUDA_CHECK(cudaMalloc(&d_S, (nw + 4) * 4));
	genome_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, ctx->stream>>>(d_S, nw, total_len, seed);
	MMB_CUDA_CHECK(cudaGetLastError());
	mi->S = (uint32_t*)calloc(nw + 1, 4);
	MMB_CUDA_CHECK(cudaMemcpyAsync(mi->S, d_S, nw * 4, cudaMemcpyDeviceToHost, ctx->stream));
	MMB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
	MMB_CUDA_CHECK(cudaFree(d_S));
	mm_idx_index_name(mi);
	mmb_idx_build_device(mi);
	return mi;
}

// Generate n_reads reads of read_len bases (ASCII, contiguous, no terminators) into out (host buffer of n_reads*read_len).
extern "C" int mmb_synth_reads(const mm_idx_t *mi, int n_reads, int read_len, uint64_t seed, float err, float sub, float ins, char *out)
{
	mm_idx_bucket_s *B = mi->B;
	mmb_ctx_t *ctx = B->ctx;
	MMB_CUDA_CHECK(cudaSetDevice(ctx->device));
	ReadGenArgs A;
	uint64_t total = 0;
	for (uint32_t i = 0; i < mi->n_seq; ++i) total += mi->seq[i].len;
	A.S = B->d_S, A.seq_off = B->d_seq_off, A.seq_len = B->d_seq_len, A.n_seq = mi->n_seq, A.total = total;
	A.n_reads = n_reads, A.read_len = read_len, A.seed = seed, A.err = err, A.sub = sub, A.ins = ins;
	char *d_out;
	MMB_CUDA_CHECK(cudaMalloc(&d_out, (size_t)n_reads * read_len + 16));
	A.out = d_out;
	reads_kernel<<<(n_reads + 127) / 128, 128, 0, ctx->stream>>>(A);
	MMB_CUDA_CHECK(cudaGetLastError());
	MMB_CUDA_CHECK(cudaMemcpyAsync(out, d_out, (size_t)n_reads * read_len, cudaMemcpyDeviceToHost, ctx->stream));
	MMB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
	MMB_CUDA_CHECK(cudaFree(d_out));
	return 0;
}

// Write the index's sequences as FASTA (80 columns are not needed by either mapper: one line per contig).
extern "C" int mmb_idx_write_fasta(const mm_idx_t *mi, const char *fn)
{
	FILE *fp = fopen(fn, "wb");
	if (!fp) return -1;
	std::string buf;
	for (uint32_t i = 0; i < mi->n_seq; ++i) {
		fprintf(fp, ">%s\n", mi->seq[i].name);
		const uint64_t off = mi->seq[i].offset;
		const uint32_t len = mi->seq[i].len;
		const uint32_t CH = 1 << 24;
		for (uint32_t b = 0; b < len; b += CH) {
			uint32_t m = len - b < CH? len - b : CH;
			buf.resize(m);
			for (uint32_t j = 0; j < m; ++j) buf[j] = "ACGTN"[mmx_seq4_get(mi->S, off + b + j) > 3? 4 : mmx_seq4_get(mi->S, off + b + j)];
			fwrite(buf.data(), 1, m, fp);
		}
		fputc('\n', fp);
	}
	fclose(fp);
	return 0;
}

// cDNA reads for the spliced workload (see cdna_reads_kernel); same output layout as mmb_synth_reads.
extern "C" int mmb_synth_cdna_reads(const mm_idx_t *mi, int n_reads, int read_len, uint64_t seed, float err, char *out)
{
	mm_idx_bucket_s *B = mi->B;
	mmb_ctx_t *ctx = B->ctx;
	MMB_CUDA_CHECK(cudaSetDevice(ctx->device));
	ReadGenArgs A;
	uint64_t total = 0;
	for (uint32_t i = 0; i < mi->n_seq; ++i) total += mi->seq[i].len;
	A.S = B->d_S, A.seq_off = B->d_seq_off, A.seq_len = B->d_seq_len, A.n_seq = mi->n_seq, A.total = total;
	A.n_reads = n_reads, A.read_len = read_len, A.seed = seed, A.err = err, A.sub = 0, A.ins = 0;
	char *d_out;
	MMB_CUDA_CHECK(cudaMalloc(&d_out, (size_t)n_reads * read_len + 16));
	A.out = d_out;
	cdna_reads_kernel<<<(n_reads + 127) / 128, 128, 0, ctx->stream>>>(A);
	MMB_CUDA_CHECK(cudaGetLastError());
	MMB_CUDA_CHECK(cudaMemcpyAsync(out, d_out, (size_t)n_reads * read_len, cudaMemcpyDeviceToHost, ctx->stream));
	MMB_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
	MMB_CUDA_CHECK(cudaFree(d_out));
	return 0;
}

// sum of qe-qs over primary hits of a batch result (the bench's "aligned bases"; all_records: every hit, for all-vs-all overlap where no
// hit is marked primary); n_regs / regs as mm_map_batch returns them
extern "C" int64_t mmb_aligned_bases(int n, const int32_t *n_regs, mm_reg1_t *const *regs, int all_records)
{
	int64_t tot = 0;
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n_regs[i]; ++j)
			if (all_records || regs[i][j].id == regs[i][j].parent) tot += regs[i][j].qe - regs[i][j].qs;
	return tot;
}
