// fmx_cub.cuh — thin wrappers over CUB device primitives (scan / radix sort / segmented sort) used by the
// index-construction, locate and regex paths.  Library code, like cuBLAS would be for a GEMM: none of it is
// on the backward-search hot loop.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fmx_kernels.cuh"

namespace fmx {

cudaError_t exclusive_sum_u32(const uint32_t *d_in, uint32_t *d_out, int64_t n, cudaStream_t st);
cudaError_t exclusive_sum_i64(const int64_t *d_in, int64_t *d_out, int64_t n, cudaStream_t st);
cudaError_t inclusive_max_u32(const uint32_t *d_in, uint32_t *d_out, int64_t n, cudaStream_t st);
// stable sort of byte keys on one bit: zeros first (the wavelet-matrix level permutation)
cudaError_t stable_partition_bit_u8(const uint8_t *d_in, uint8_t *d_out, int64_t n, int bit, cudaStream_t st);
// stable sort of byte keys on the digit [begin_bit, end_bit): the multi-ary wavelet-matrix level permutation
cudaError_t stable_partition_digit_u8(const uint8_t *d_in, uint8_t *d_out, int64_t n, int begin_bit, int end_bit, cudaStream_t st);
cudaError_t sort_pairs_u64_u32(const uint64_t *k_in, uint64_t *k_out, const uint32_t *v_in, uint32_t *v_out, int64_t n,
                               int begin_bit, int end_bit, cudaStream_t st);
cudaError_t sort_pairs_u8_u32(const uint8_t *k_in, uint8_t *k_out, const uint32_t *v_in, uint32_t *v_out, int64_t n,
                              cudaStream_t st);
// ascending sort of 64-bit keys on bits [0, end_bit): the (query, position) keys of a locate slab
cudaError_t radix_sort_u64(const uint64_t *k_in, uint64_t *k_out, int64_t n, int end_bit, cudaStream_t st);
// plain ascending sort (one segment of any size)
cudaError_t radix_sort_u32(const uint32_t *k_in, uint32_t *k_out, int64_t n, cudaStream_t st);
cudaError_t widen_u32_i64(const uint32_t *d_in, int64_t *d_out, int64_t n, cudaStream_t st);
// sort regex results by (regex, len, sp, ep); d_tmp is scratch of the same size
cudaError_t sort_regex_results(RegexResult *d_res, RegexResult *d_tmp, int64_t n, uint32_t n_regex, uint32_t max_len, cudaStream_t st);

}  // namespace fmx
