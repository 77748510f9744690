// tools/ldhint_bench.cu — which load form gives the most random requests per second on B200, and how many DRAM bytes
// does each request really move?  Random aligned gathers of 4/16/32/64 B over a large buffer, one variant per launch:
//   0  ld.global.nc.L1::no_allocate            (what the kernels used in round 1)
//   1  ... + .L2::64B      2  ... + .L2::128B      3  ... + .L2::256B      (prefetch-size hints)
//   4  ld.global.cg        5  ld.global.cv        6  ld.global (default, L1 allocate)
//   7  ld.global.nc.L1::no_allocate.v8.u32  (256-bit load, 32 B per lane: gathers of 32/64/128 B by 1/2/4 lanes)
//   8  the same with .L2::64B      9  whole chunk by one thread (several loads) / two lanes
// Run plain for rates; run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum` for bytes per request.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/ldhint_bench tools/ldhint_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

template <int V> __device__ __forceinline__ uint4 ld16(const uint4 *p) {
    uint4 r;
    if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 4) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 5) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 6) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    if (V == 7) {                                   // sm_100 256-bit load: this lane fetches 32 B (p counts 32-B units for this variant)
        uint32_t a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w), "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        r.x ^= a; r.y ^= b; r.z ^= c; r.w ^= d;
    }
    if (V == 8) {                                   // 256-bit load with the 64-B L2 prefetch-size hint
        uint32_t a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w), "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        r.x ^= a; r.y ^= b; r.z ^= c; r.w ^= d;
    }
    return r;
}
template <int V> __device__ __forceinline__ uint32_t ld4(const uint32_t *p) {
    uint32_t r;
    if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 4) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 5) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 6) asm volatile("ld.global.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 7) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    if (V == 8) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// LANES lanes fetch one aligned chunk of LANES*16 bytes (LANES = 0 means one 4-byte load by one lane); `chain` dependent gathers each
template <int V, int LANES, int VEC = 1>
__global__ void __launch_bounds__(256) gather(const uint4 *base, unsigned long long units, long long gathers, int chain, uint32_t seed,
                                              unsigned long long *sink) {
    constexpr int L = LANES == 0 ? 1 : LANES;
    const long long gid = ((long long)blockIdx.x * 256 + threadIdx.x) / L;
    const int lane = threadIdx.x % L;
    if (gid >= gathers) return;
    unsigned long long x = (unsigned long long)gid * 0x9E3779B97F4A7C15ull + seed;
    uint32_t acc = 0;
    for (int s = 0; s < chain; ++s) {
        x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
        const unsigned long long unit = __umul64hi(x, units);
        uint32_t v;
        if (LANES == 0) v = ld4<V>(reinterpret_cast<const uint32_t *>(base) + unit);
        else {
            v = 0;
#pragma unroll
            for (int i = 0; i < VEC; ++i) { const uint4 q = ld16<V>(base + ((unit * L + lane) * VEC + i) * (V >= 7 ? 2 : 1)); v ^= q.x ^ q.y ^ q.z ^ q.w; }
        }
        if (L > 1) {
            const uint32_t mask = ((1u << L) - 1u) << ((threadIdx.x & 31) & ~(L - 1));
            for (int o = 1; o < L; o <<= 1) v ^= __shfl_xor_sync(mask, v, o);
        }
        acc ^= v;
        x += v;
    }
    if (acc == 0x12345678u && lane == 0) atomicAdd(sink, 1ull);
}

template <int V, int LANES, int VEC = 1>
static void run(const char *vname, const uint4 *buf, size_t bytes, long long gathers, int chain, int iters, unsigned long long *sink) {
    const int L = LANES == 0 ? 1 : LANES;
    const int gbytes = LANES == 0 ? 4 : LANES * VEC * (V >= 7 ? 32 : 16);
    const unsigned long long units = bytes / gbytes;
    const unsigned grid = (unsigned)((gathers * L + 255) / 256);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather<V, LANES, VEC><<<grid, 256>>>(buf, units, gathers, chain, 1u, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < iters; ++it) {
        CK(cudaEventRecord(e0));
        gather<V, LANES, VEC><<<grid, 256>>>(buf, units, gathers, chain, 100u + it, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double req = (double)gathers * chain;
    printf("{\"variant\": \"%s\", \"v\": %d, \"gather_bytes\": %d, \"lanes\": %d, \"loads_per_lane\": %d, \"footprint_gb\": %.1f, \"chain\": %d, \"ms\": %.4f, \"greq_per_s\": %.2f, \"useful_gbs\": %.1f}\n",
           vname, V, gbytes, L, VEC, bytes / 1e9, chain, best, req / best / 1e6, req * gbytes / best / 1e6);
    fflush(stdout);
}

template <int V> static void run_all(const char *vname, const uint4 *buf, size_t bytes, long long gathers, int chain, int iters, unsigned long long *sink, int only) {
    if (only < 0 || only == 4) run<V, 0>(vname, buf, bytes, gathers, chain, iters, sink);
    if (only < 0 || only == 16) run<V, 1>(vname, buf, bytes, gathers, chain, iters, sink);
    if (only < 0 || only == 32) run<V, 2>(vname, buf, bytes, gathers, chain, iters, sink);
    if (only < 0 || only == 64) run<V, 4>(vname, buf, bytes, gathers, chain, iters, sink);
}

__global__ void fill(uint4 *p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4((uint32_t)i * 2654435761u, (uint32_t)(i >> 7), (uint32_t)i ^ 0x5bd1e995u, (uint32_t)(i * 40503u));
}

int main(int argc, char **argv) {
    const double gb = argc > 1 ? atof(argv[1]) : 34.0;
    const long long gathers = argc > 2 ? atoll(argv[2]) : (1ll << 24);
    const int chain = argc > 3 ? atoi(argv[3]) : 8;
    const int iters = argc > 4 ? atoi(argv[4]) : 3;
    const int only = argc > 5 ? atoi(argv[5]) : -1;          // gather size filter
    const int onlyv = argc > 6 ? atoi(argv[6]) : -1;         // variant filter
    const size_t bytes = (size_t)(gb * 1e9) / 256 * 256;
    uint4 *buf; unsigned long long *sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    fill<<<148 * 8, 256>>>(buf, bytes / 16);
    CK(cudaDeviceSynchronize());
#define RV(V, NAME) if (onlyv < 0 || onlyv == V) run_all<V>(NAME, buf, bytes, gathers, chain, iters, sink, only)
    RV(0, "nc.L1::no_allocate");
    RV(1, "nc.L1::no_allocate.L2::64B");
    RV(2, "nc.L1::no_allocate.L2::128B");
    RV(3, "nc.L1::no_allocate.L2::256B");
    RV(4, "cg");
    RV(5, "cv");
    RV(6, "default");
    RV(7, "nc.L1::no_allocate.v8 (32 B per lane)");
    RV(8, "nc.L1::no_allocate.L2::64B.v8 (32 B per lane)");
    if (onlyv < 0 || onlyv == 9) {                  // one thread fetches the whole chunk with several loads
        run<8, 1, 2>("1 lane x 2 x v8.L2::64B", buf, bytes, gathers, chain, iters, sink);
        run<1, 1, 4>("1 lane x 4 x v4.L2::64B", buf, bytes, gathers, chain, iters, sink);
        run<7, 1, 4>("1 lane x 4 x v8", buf, bytes, gathers, chain, iters, sink);
        run<7, 2, 2>("2 lanes x 2 x v8", buf, bytes, gathers, chain, iters, sink);
        run<1, 2, 2>("2 lanes x 2 x v4.L2::64B", buf, bytes, gathers, chain, iters, sink);
    }
    return 0;
}
