// fmx_cub.cu — see fmx_cub.cuh
#include "fmx_cub.cuh"

#include <cub/cub.cuh>

namespace fmx {

namespace {
struct Temp {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t st) { return cudaMallocAsync(&p, bytes ? bytes : 1, st); }
    void release(cudaStream_t st) { if (p) cudaFreeAsync(p, st); p = nullptr; }
};
struct MaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };
}  // namespace

#define FMX_CUB2(EXPR_SIZE, EXPR_RUN)                            \
    do {                                                         \
        size_t bytes = 0; void *tp = nullptr;                    \
        cudaError_t e = (EXPR_SIZE);                             \
        if (e != cudaSuccess) return e;                          \
        Temp t; e = t.alloc(bytes, st);                          \
        if (e != cudaSuccess) return e;                          \
        tp = t.p;                                                \
        e = (EXPR_RUN);                                          \
        t.release(st);                                           \
        return e;                                                \
    } while (0)

cudaError_t exclusive_sum_u32(const uint32_t *d_in, uint32_t *d_out, int64_t n, cudaStream_t st) {
    FMX_CUB2(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in, d_out, n, st),
             cub::DeviceScan::ExclusiveSum(tp, bytes, d_in, d_out, n, st));
}
cudaError_t exclusive_sum_i64(const int64_t *d_in, int64_t *d_out, int64_t n, cudaStream_t st) {
    FMX_CUB2(cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in, d_out, n, st),
             cub::DeviceScan::ExclusiveSum(tp, bytes, d_in, d_out, n, st));
}
cudaError_t inclusive_max_u32(const uint32_t *d_in, uint32_t *d_out, int64_t n, cudaStream_t st) {
    FMX_CUB2(cub::DeviceScan::InclusiveScan(nullptr, bytes, d_in, d_out, MaxOp(), n, st),
             cub::DeviceScan::InclusiveScan(tp, bytes, d_in, d_out, MaxOp(), n, st));
}
cudaError_t stable_partition_bit_u8(const uint8_t *d_in, uint8_t *d_out, int64_t n, int bit, cudaStream_t st) {
    FMX_CUB2(cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_in, d_out, n, bit, bit + 1, st),
             cub::DeviceRadixSort::SortKeys(tp, bytes, d_in, d_out, n, bit, bit + 1, st));
}
cudaError_t stable_partition_digit_u8(const uint8_t *d_in, uint8_t *d_out, int64_t n, int begin_bit, int end_bit, cudaStream_t st) {
    FMX_CUB2(cub::DeviceRadixSort::SortKeys(nullptr, bytes, d_in, d_out, n, begin_bit, end_bit, st),
             cub::DeviceRadixSort::SortKeys(tp, bytes, d_in, d_out, n, begin_bit, end_bit, st));
}
cudaError_t sort_pairs_u64_u32(const uint64_t *k_in, uint64_t *k_out, const uint32_t *v_in, uint32_t *v_out, int64_t n,
                               int begin_bit, int end_bit, cudaStream_t st) {
    FMX_CUB2(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, n, begin_bit, end_bit, st),
             cub::DeviceRadixSort::SortPairs(tp, bytes, k_in, k_out, v_in, v_out, n, begin_bit, end_bit, st));
}
cudaError_t sort_pairs_u8_u32(const uint8_t *k_in, uint8_t *k_out, const uint32_t *v_in, uint32_t *v_out, int64_t n,
                              cudaStream_t st) {
    FMX_CUB2(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, n, 0, 8, st),
             cub::DeviceRadixSort::SortPairs(tp, bytes, k_in, k_out, v_in, v_out, n, 0, 8, st));
}
cudaError_t radix_sort_u64(const uint64_t *k_in, uint64_t *k_out, int64_t n, int end_bit, cudaStream_t st) {
    FMX_CUB2(cub::DeviceRadixSort::SortKeys(nullptr, bytes, k_in, k_out, n, 0, end_bit, st),
             cub::DeviceRadixSort::SortKeys(tp, bytes, k_in, k_out, n, 0, end_bit, st));
}
cudaError_t radix_sort_u32(const uint32_t *k_in, uint32_t *k_out, int64_t n, cudaStream_t st) {
    FMX_CUB2(cub::DeviceRadixSort::SortKeys(nullptr, bytes, k_in, k_out, n, 0, 32, st),
             cub::DeviceRadixSort::SortKeys(tp, bytes, k_in, k_out, n, 0, 32, st));
}
// ---- regex result ordering: LSD radix over the two 64-bit halves of (regex,len | sp,ep) ------------------
__global__ void split_keys_kernel(const RegexResult *r, int64_t n, uint64_t *hi, uint64_t *lo, uint32_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RegexResult v = r[i];
    hi[i] = ((uint64_t)v.regex << 32) | v.len;
    lo[i] = ((uint64_t)v.sp << 32) | v.ep;
    idx[i] = (uint32_t)i;
}
__global__ void gather_u64_kernel(const uint64_t *src, const uint32_t *idx, int64_t n, uint64_t *dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void gather_res_kernel(const RegexResult *src, const uint32_t *idx, int64_t n, RegexResult *dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void widen_kernel(const uint32_t *__restrict__ in, long long *__restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (long long)in[i];
}
cudaError_t widen_u32_i64(const uint32_t *d_in, int64_t *d_out, int64_t n, cudaStream_t st) {
    if (n > 0) widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, (long long *)d_out, n);
    return cudaGetLastError();
}

// one-key form: (regex, len, sp) packed into 64 bits orders the results completely — two results of one regex with the same length and
// the same sp are the same string, hence the same interval — and only the used bits are sorted
__global__ void pack_keys_kernel(const RegexResult *r, int64_t n, int len_bits, uint64_t *key, uint32_t *idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RegexResult v = r[i];
    key[i] = ((((uint64_t)v.regex << len_bits) | v.len) << 32) | v.sp;
    idx[i] = (uint32_t)i;
}
cudaError_t sort_regex_results(RegexResult *d_res, RegexResult *d_tmp, int64_t n, uint32_t n_regex, uint32_t max_len, cudaStream_t st) {
    if (n <= 1) return cudaSuccess;
    if (n >= (1ll << 32)) return cudaErrorInvalidValue;
    int rb = 1, lb = 1;
    while ((1ull << rb) < (uint64_t)n_regex) ++rb;
    while ((1ull << lb) <= (uint64_t)max_len) ++lb;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaError_t e;
    if (rb + lb <= 32) {
        uint64_t *k0, *k1; uint32_t *i0, *i1;
        if ((e = cudaMallocAsync(&k0, n * 8, st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(&k1, n * 8, st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(&i0, n * 4, st)) != cudaSuccess) return e;
        if ((e = cudaMallocAsync(&i1, n * 4, st)) != cudaSuccess) return e;
        pack_keys_kernel<<<grid, 256, 0, st>>>(d_res, n, lb, k0, i0);
        e = sort_pairs_u64_u32(k0, k1, i0, i1, n, 0, 32 + rb + lb, st);
        if (e == cudaSuccess) {
            gather_res_kernel<<<grid, 256, 0, st>>>(d_res, i1, n, d_tmp);
            e = cudaMemcpyAsync(d_res, d_tmp, n * sizeof(RegexResult), cudaMemcpyDeviceToDevice, st);
        }
        cudaFreeAsync(k0, st); cudaFreeAsync(k1, st); cudaFreeAsync(i0, st); cudaFreeAsync(i1, st);
        return e;
    }
    uint64_t *hi, *lo, *k2; uint32_t *i0, *i1;
    if ((e = cudaMallocAsync(&hi, n * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&lo, n * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&k2, n * 8, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&i0, n * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMallocAsync(&i1, n * 4, st)) != cudaSuccess) return e;
    split_keys_kernel<<<grid, 256, 0, st>>>(d_res, n, hi, lo, i0);
    e = sort_pairs_u64_u32(lo, k2, i0, i1, n, 0, 64, st);                 // by (sp,ep)
    if (e == cudaSuccess) {
        gather_u64_kernel<<<grid, 256, 0, st>>>(hi, i1, n, lo);           // lo := hi permuted
        e = sort_pairs_u64_u32(lo, k2, i1, i0, n, 0, 64, st);             // stable by (regex,len)
    }
    if (e == cudaSuccess) {
        gather_res_kernel<<<grid, 256, 0, st>>>(d_res, i0, n, d_tmp);
        e = cudaMemcpyAsync(d_res, d_tmp, n * sizeof(RegexResult), cudaMemcpyDeviceToDevice, st);
    }
    cudaFreeAsync(hi, st); cudaFreeAsync(lo, st); cudaFreeAsync(k2, st); cudaFreeAsync(i0, st); cudaFreeAsync(i1, st);
    return e;
}

}  // namespace fmx
