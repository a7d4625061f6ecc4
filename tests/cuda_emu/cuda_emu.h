// tests/cuda_emu/cuda_emu.h -- TEST INFRASTRUCTURE: a small SIMT emulator that lets the product's CUDA kernel sources run on the CPU
// (one OS thread per CUDA thread, warp/block collectives as rendezvous), so that kernel LOGIC can be checked against the oracle in
// the CPU test suite. It says nothing about performance, memory coalescing or the GPU memory model. Force-included
// (g++ -include) in front of the transformed kernel sources produced by tests/cuda_emu/build_emu.py.
#pragma once
#include <cuda_runtime.h> // types and API declarations only; the functions are provided by emu_runtime.cc
#include <cstdint>
#include <cstring>
#include <functional>

#undef __shared__
#define __shared__ static          /* blocks run one after another, so block-shared == process-static */
#undef __launch_bounds__
#define __launch_bounds__(...)
#define MMB_EMU 1
#define cudaFuncSetAttribute(...) (cudaSuccess)                      /* kernels are plain functions here */
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, ...) (*(n) = 1, cudaSuccess)

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
void *emu_dyn_smem();
void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
void emu_trace(const char *kernel_name); // EMU_TRACE=1: print every launch to stderr (debugging aid)

// ---- warp / block collectives (emu_runtime.cc) ----
void emu_warp_exchange(unsigned mask, uint64_t mine, uint64_t out[32], unsigned *present); // rendezvous of the live lanes in mask
void __syncthreads();
static inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t o[32]; unsigned p; emu_warp_exchange(mask, 0, o, &p); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int emu_lane() { return (int)((threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x) & 31; }

template<class T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template<class T> static inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template<class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	uint64_t o[32]; unsigned p; const int lane = emu_lane();
	emu_warp_exchange(mask, emu_bits(v), o, &p);
	const int s = (lane & ~(width - 1)) | (src & (width - 1));
	return (p >> s & 1)? emu_unbits<T>(o[s]) : v;
}
template<class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	uint64_t o[32]; unsigned p; const int lane = emu_lane();
	emu_warp_exchange(mask, emu_bits(v), o, &p);
	const int s = lane - (int)delta;
	if (s < (lane & ~(width - 1))) return v;
	return (p >> s & 1)? emu_unbits<T>(o[s]) : v;
}
template<class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	uint64_t o[32]; unsigned p; const int lane = emu_lane();
	emu_warp_exchange(mask, emu_bits(v), o, &p);
	const int s = lane + (int)delta;
	if (s >= (lane & ~(width - 1)) + width) return v;
	return (p >> s & 1)? emu_unbits<T>(o[s]) : v;
}
template<class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lm, int width = 32)
{
	uint64_t o[32]; unsigned p; const int lane = emu_lane();
	emu_warp_exchange(mask, emu_bits(v), o, &p);
	const int s = lane ^ lm;
	if (s >= (lane & ~(width - 1)) + width || s < (lane & ~(width - 1))) return v;
	return (p >> s & 1)? emu_unbits<T>(o[s]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	uint64_t o[32]; unsigned p, r = 0;
	emu_warp_exchange(mask, pred? 1 : 0, o, &p);
	for (int l = 0; l < 32; ++l) if ((p >> l & 1) && o[l]) r |= 1u << l;
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { uint64_t o[32]; unsigned p; emu_warp_exchange(mask, pred? 1 : 0, o, &p); for (int l = 0; l < 32; ++l) if ((p >> l & 1) && !o[l]) return 0; return 1; }
#define EMU_REDUCE(NAME, T, INIT, OP) static inline T NAME(unsigned mask, T v) { uint64_t o[32]; unsigned p; emu_warp_exchange(mask, emu_bits(v), o, &p); \
	bool first = true; T r = INIT; for (int l = 0; l < 32; ++l) if (p >> l & 1) { T x = emu_unbits<T>(o[l]); if (first) r = x, first = false; else r = OP; } return r; }
EMU_REDUCE(__reduce_max_sync, int, 0, (x > r? x : r))
EMU_REDUCE(__reduce_min_sync, int, 0, (x < r? x : r))
EMU_REDUCE(__reduce_add_sync, int, 0, (r + x))
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, (r | x))
EMU_REDUCE(__reduce_and_sync, unsigned, 0u, (r & x))

// ---- atomics ----
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline int atomicMax(int *p, int v) { int o = *p; while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

// ---- min/max as CUDA provides them in the global namespace ----
#include <algorithm>
static inline int min(int a, int b) { return a < b? a : b; }
static inline int max(int a, int b) { return a > b? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b? a : b; }
static inline long long min(long long a, long long b) { return a < b? a : b; }
static inline long long max(long long a, long long b) { return a > b? a : b; }
static inline float min(float a, float b) { return a < b? a : b; }
static inline float max(float a, float b) { return a > b? a : b; }

// ---- integer intrinsics ----
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned __match_any_sync(unsigned mask, unsigned v)
{
	uint64_t o[32]; unsigned p;
	emu_warp_exchange(mask, v, o, &p);
	unsigned r = 0;
	for (int l = 0; l < 32; ++l) if ((p >> l & 1) && (unsigned)o[l] == v) r |= 1u << l;
	return r;
}
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= (x >> i & 1u) << (31 - i); return r; }
static inline unsigned long long __brevll(unsigned long long x) { unsigned long long r = 0; for (int i = 0; i < 64; ++i) r |= (x >> i & 1ull) << (63 - i); return r; }
static inline int __clz(int x) { return x == 0? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { const uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)(v >> (sh & 31)); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
	const uint64_t v = (uint64_t)y << 32 | x;
	unsigned r = 0;
	for (int i = 0; i < 4; ++i) {
		const unsigned sel = s >> (4 * i) & 0xf;
		unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
		if (sel & 8) b = (b & 0x80)? 0xff : 0;
		r |= b << (8 * i);
	}
	return r;
}
#define EMU_V4(NAME, EXPR) static inline unsigned NAME(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; ++i) { const int8_t x = (int8_t)(a >> (8 * i)), y = (int8_t)(b >> (8 * i)); \
	const uint8_t ux = (uint8_t)x, uy = (uint8_t)y; (void)ux; (void)uy; r |= (unsigned)(uint8_t)(EXPR) << (8 * i); } return r; }
EMU_V4(__vadd4, ux + uy)
EMU_V4(__vsub4, ux - uy)
EMU_V4(__vcmpeq4, ux == uy? 0xff : 0)
EMU_V4(__vcmpgts4, x > y? 0xff : 0)
EMU_V4(__vmins4, x < y? x : y)
EMU_V4(__vmaxs4, x > y? x : y)
#define EMU_V2(NAME, EXPR) static inline unsigned NAME(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 2; ++i) { const int16_t x = (int16_t)(a >> (16 * i)), y = (int16_t)(b >> (16 * i)); \
	const uint16_t ux = (uint16_t)x, uy = (uint16_t)y; (void)ux; (void)uy; (void)x; (void)y; r |= (unsigned)(uint16_t)(EXPR) << (16 * i); } return r; }
EMU_V2(__vadd2, ux + uy)
EMU_V2(__vsub2, ux - uy)
EMU_V2(__vmaxs2, x > y? x : y)
EMU_V2(__vmins2, x < y? x : y)
EMU_V2(__vmaxu2, ux > uy? ux : uy)
EMU_V2(__vminu2, ux < uy? ux : uy)
static inline unsigned __vimax3_s16x2(unsigned a, unsigned b, unsigned c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
