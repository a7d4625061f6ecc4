// tests/cuda_emu/emu_runtime.cc -- TEST INFRASTRUCTURE (see cuda_emu.h): thread-per-CUDA-thread launcher, warp/block rendezvous, and
// a host-memory implementation of the handful of CUDA runtime calls the product's host code makes.
#include "cuda_emu.h"
#include <thread>
#include <vector>
#include <mutex>
#include <condition_variable>
#include <map>
#include <cstdio>
#include <cstdlib>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace {
struct Rendezvous { unsigned arrived = 0; uint64_t gen = 0; uint64_t val[2][32]; unsigned present[2] = {0, 0}; };
struct WarpState { std::mutex mu; std::condition_variable cv; unsigned alive = 0; std::map<unsigned, Rendezvous> rv; };
struct BlockState {
	std::vector<WarpState> warps;
	std::mutex mu; std::condition_variable cv;
	int alive = 0, arrived = 0; uint64_t gen = 0;
	std::vector<uint8_t> smem;
};
BlockState *g_block = nullptr;
thread_local int tl_tid = 0;
}

void *emu_dyn_smem() { return g_block->smem.data(); }

// All live lanes named in `mask` meet here; every participant gets everyone's value. Lanes that already returned from the kernel
// are not waited for (the hardware does not wait for exited threads either).
void emu_warp_exchange(unsigned mask, uint64_t mine, uint64_t out[32], unsigned *present)
{
	WarpState &W = g_block->warps[tl_tid >> 5];
	const int lane = tl_tid & 31;
	std::unique_lock<std::mutex> lk(W.mu);
	Rendezvous &R = W.rv[mask];
	const uint64_t my_gen = R.gen;
	const int par = (int)(my_gen & 1);
	R.val[par][lane] = mine, R.present[par] |= 1u << lane, R.arrived |= 1u << lane;
	auto complete = [&]() { return (R.arrived & W.alive & mask) == (W.alive & mask); };
	if (complete()) {
		R.arrived = 0, ++R.gen;
		R.present[1 - par] = 0; // the buffer of the next generation starts empty
		W.cv.notify_all();
	} else W.cv.wait(lk, [&]() { if (R.gen != my_gen) return true; if (complete()) { R.arrived = 0, ++R.gen; R.present[1 - par] = 0; W.cv.notify_all(); return true; } return false; });
	memcpy(out, R.val[par], sizeof(uint64_t) * 32);
	*present = R.present[par];
}

void __syncthreads()
{
	BlockState &B = *g_block;
	std::unique_lock<std::mutex> lk(B.mu);
	const uint64_t g = B.gen;
	if (++B.arrived >= B.alive) { B.arrived = 0, ++B.gen; B.cv.notify_all(); }
	else B.cv.wait(lk, [&]() { return B.gen != g; });
}

// One set of OS threads per launch; the blocks of the grid run one after another on it (thread t plays CUDA thread t of every block).
void emu_trace(const char *kernel_name)
{
	static const bool on = getenv("EMU_TRACE") != nullptr;
	if (on) fprintf(stderr, "[emu] launch %s\n", kernel_name);
}

void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body)
{
	const int nt = (int)(block.x * block.y * block.z);
	const long n_blocks = (long)grid.x * grid.y * grid.z;
	if (n_blocks <= 0 || nt <= 0) return;
	BlockState B;
	std::mutex gmu; std::condition_variable gcv; // block-to-block hand-over
	long cur = 0; int done = 0;
	auto reset_block = [&]() {
		B.warps = std::vector<WarpState>((nt + 31) / 32);
		for (int t = 0; t < nt; ++t) B.warps[t >> 5].alive |= 1u << (t & 31);
		B.alive = nt, B.arrived = 0, B.gen = 0;
		B.smem.assign(smem_bytes + 64, 0);
	};
	reset_block();
	g_block = &B;
	std::vector<std::thread> th;
	for (int t = 0; t < nt; ++t)
		th.emplace_back([&, t]() {
			tl_tid = t;
			blockDim = block, gridDim = grid;
			threadIdx.x = t % block.x, threadIdx.y = t / block.x % block.y, threadIdx.z = t / (block.x * block.y);
			for (long bi = 0; bi < n_blocks; ++bi) {
				blockIdx.x = (unsigned)(bi % grid.x), blockIdx.y = (unsigned)(bi / grid.x % grid.y), blockIdx.z = (unsigned)(bi / ((long)grid.x * grid.y));
				body();
				{ // the thread leaves the block: pending rendezvous of its warp / block may now be complete
					WarpState &W = B.warps[t >> 5];
					std::lock_guard<std::mutex> lk(W.mu);
					W.alive &= ~(1u << (t & 31));
					W.cv.notify_all();
				}
				{
					std::lock_guard<std::mutex> lk(B.mu);
					--B.alive;
					if (B.alive > 0 && B.arrived >= B.alive) { B.arrived = 0, ++B.gen; }
					B.cv.notify_all();
				}
				// wait until every thread has left, then the last one prepares the next block
				std::unique_lock<std::mutex> lk(gmu);
				if (++done == nt) { done = 0; if (bi + 1 < n_blocks) reset_block(); ++cur; gcv.notify_all(); }
				else gcv.wait(lk, [&]() { return cur > bi; });
			}
		});
	for (auto &x : th) x.join();
	g_block = nullptr;
}

// ---- CUDA runtime on host memory ----
extern "C" {
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); p->multiProcessorCount = 2; p->sharedMemPerBlockOptin = 232448; p->major = 10; strcpy(p->name, "SIMT emulator"); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n + 64, 1); return *p? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n + 64, 1); return *p? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
}
