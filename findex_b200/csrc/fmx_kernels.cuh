// fmx_kernels.cuh — host-callable launchers of the search kernels (definitions in fmx_kernels.cu).
#pragma once
#include "fmx_device.cuh"

namespace fmx {

constexpr int kThreads = 256;

struct LaunchCfg {
    int layout;          // FMX_LAYOUT_WM / FMX_LAYOUT_PLANES
    int lanes;           // 1, 2 or 4 lanes per query
    int count_lanes = 0;       // > 0: lanes per query of the count kernels only (one lane once row contexts answer most queries in a
                               // single 32-byte request each: twice the queries in flight, half the instructions per query)
    int min_blocks = 0;        // 4 or 6: the device-pointer count kernel compiled for that many resident CTAs per SM (0 = default: 8, 6 at one lane)
};

// fused exchange: up to 8 gathered buffers (one per rank, peer-mapped) that receive this shard's hit counts
struct PeerSinks {
    uint32_t *p[8];
    int32_t   n;
    long long offset;            // element offset of this shard inside each gathered buffer
};

// count: fixed-length patterns (device pointers).  out64 selects int64 vs uint32 outputs.
cudaError_t launch_count_fixed(const DevIndex &ix, LaunchCfg cfg, const uint8_t *d_pat, int len, int64_t m,
                               void *d_sp, void *d_ep, bool out64, unsigned long long *d_stats, cudaStream_t st,
                               const PeerSinks *sinks_or_null = nullptr);
// count: variable-length patterns with int64 offsets.
cudaError_t launch_count_var(const DevIndex &ix, LaunchCfg cfg, const uint8_t *d_pat, const int64_t *d_off, int64_t m,
                             int64_t *d_sp, int64_t *d_ep, cudaStream_t st);
cudaError_t launch_occ(const DevIndex &ix, LaunchCfg cfg, const uint8_t *d_c, const int64_t *d_key, int64_t m,
                       int64_t *d_out, cudaStream_t st);
cudaError_t launch_prev_range(const DevIndex &ix, LaunchCfg cfg, const int64_t *d_sp, const int64_t *d_ep,
                              const uint8_t *d_c, int64_t m, int64_t *d_sp1, int64_t *d_ep1, cudaStream_t st);
// getIntervalPrevRange: one interval, chars cstart..cend; outputs indexed by (c - cstart)
cudaError_t launch_interval_prev_range(const DevIndex &ix, LaunchCfg cfg, int64_t sp, int64_t ep, int cstart, int cend,
                                       int64_t *d_sp1, int64_t *d_ep1, cudaStream_t st);
// LF step for rows (getPrevI)
cudaError_t launch_lf(const DevIndex &ix, LaunchCfg cfg, const int64_t *d_row, int64_t m, int64_t *d_out, cudaStream_t st);
// FL step for rows (getNextI) and nextSubstr
cudaError_t launch_fl(const DevIndex &ix, LaunchCfg cfg, const int64_t *d_row, int64_t m, int64_t *d_out, cudaStream_t st);
cudaError_t launch_next_substr(const DevIndex &ix, LaunchCfg cfg, const int64_t *d_row, int64_t m, int len, uint8_t *d_out,
                               int *d_out_len, cudaStream_t st);
// prevSubstr: len LF steps per row, emitting the BWT byte at each visited row
cudaError_t launch_prev_substr(const DevIndex &ix, LaunchCfg cfg, const int64_t *d_row, int64_t m, int len,
                               uint8_t *d_out, cudaStream_t st);
// locate: one work item per occurrence; d_off[m+1] exclusive offsets, writes unsorted sa values
cudaError_t launch_locate(const DevIndex &ix, LaunchCfg cfg, const int64_t *d_sp, const int64_t *d_off, int64_t m,
                          int64_t total, int sample_rate, uint32_t *d_pos, cudaStream_t st);

// regex frontier
struct RegexTables {
    const uint8_t  *st_c;        // per global state
    const uint8_t  *st_last;
    const uint32_t *st_regex;    // owning regex index in the batch
    const uint32_t *fol_off;     // CSR over global states
    const uint32_t *fol;         // global state ids
};
struct FrontierItem { uint32_t state, len, sp, ep; };
struct RegexResult  { uint32_t regex, len, sp, ep; };
// the whole traversal in one cooperative launch (persistent grid, grid-wide barrier between levels).  d_ctrl: 8 x u64, zeroed by the caller:
// [3] = matches found (may exceed cap_res: writes are dropped, the count keeps growing), [4] = status (0 ok; 1 = a frontier of [5] items
// outgrew `cap`: regrow and rerun; 2 = more than max_levels levels), [6] = levels run.
cudaError_t launch_regex_search(const DevIndex &ix, LaunchCfg cfg, RegexTables rt, const uint32_t *d_first, int64_t n_first,
                                FrontierItem *d_a, FrontierItem *d_b, int64_t cap, RegexResult *d_res, int64_t cap_res,
                                unsigned long long *d_ctrl, int64_t max_levels, cudaStream_t st);

// K4 random gather microbenchmark
cudaError_t launch_gather_bench(const uint4 *base, uint64_t n_blocks64, int bytes_per_gather, int lanes, int64_t gathers,
                                int chain, uint32_t seed, unsigned long long *d_sink, cudaStream_t st);

}  // namespace fmx
