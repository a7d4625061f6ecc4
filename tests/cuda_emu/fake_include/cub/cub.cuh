// tests/cuda_emu/fake_include/cub/cub.cuh -- TEST INFRASTRUCTURE: host stand-ins for the four CUB device-wide primitives the index
// build uses (csrc/index.cu), with CUB's two-phase calling convention (first call with a null workspace returns its size).
#pragma once
#include <algorithm>
#include <vector>
#include <cstdint>
#include <cstddef>
namespace cub {
struct DeviceRadixSort {
	template<class K> static cudaError_t SortKeys(void *tmp, size_t &bytes, const K *in, K *out, int64_t n, int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = 0)
	{
		if (!tmp) { bytes = 16; return cudaSuccess; }
		const K mask = end_bit >= (int)sizeof(K) * 8? ~(K)0 : (((K)1 << end_bit) - 1);
		std::vector<K> v(in, in + n);
		std::stable_sort(v.begin(), v.end(), [&](K a, K b) { return ((a & mask) >> begin_bit) < ((b & mask) >> begin_bit); });
		std::copy(v.begin(), v.end(), out);
		return cudaSuccess;
	}
	template<class K, class V> static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int64_t n, int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = 0)
	{
		if (!tmp) { bytes = 16; return cudaSuccess; }
		const K mask = end_bit >= (int)sizeof(K) * 8? ~(K)0 : (((K)1 << end_bit) - 1);
		std::vector<int64_t> idx(n);
		for (int64_t i = 0; i < n; ++i) idx[i] = i;
		std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return ((kin[a] & mask) >> begin_bit) < ((kin[b] & mask) >> begin_bit); });
		std::vector<K> k(n); std::vector<V> v(n);
		for (int64_t i = 0; i < n; ++i) k[i] = kin[idx[i]], v[i] = vin[idx[i]];
		std::copy(k.begin(), k.end(), kout); std::copy(v.begin(), v.end(), vout);
		return cudaSuccess;
	}
};
struct DeviceRunLengthEncode {
	template<class In, class U, class Cn, class N> static cudaError_t Encode(void *tmp, size_t &bytes, In in, U uniq, Cn cnt, N n_runs, int64_t n, cudaStream_t = 0)
	{
		if (!tmp) { bytes = 16; return cudaSuccess; }
		int64_t r = 0;
		for (int64_t i = 0; i < n;) { int64_t j = i; while (j < n && in[j] == in[i]) ++j; uniq[r] = in[i], cnt[r] = (decltype(+cnt[0]))(j - i), ++r, i = j; }
		*n_runs = (decltype(+*n_runs))r;
		return cudaSuccess;
	}
};
struct DeviceScan {
	template<class In, class Out> static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, In in, Out out, int64_t n, cudaStream_t = 0)
	{
		if (!tmp) { bytes = 16; return cudaSuccess; }
		decltype(+out[0]) run = 0;
		for (int64_t i = 0; i < n; ++i) { const auto v = in[i]; out[i] = run; run += v; }
		return cudaSuccess;
	}
};
}
