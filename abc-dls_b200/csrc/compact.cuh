// compact.cuh — ordered (stable) stream compaction building blocks: per-tile counts -> exclusive scan
// of the tile counts -> per-tile write.  Row order is preserved, which R's logical indexing
// (param[wt1, ]) and the cumsum cap rule both require.
#pragma once
#include "common.cuh"

namespace {

constexpr int kTileThreads = 256;
constexpr int kItems = 8;                      // consecutive rows per thread
constexpr int kTile = kTileThreads * kItems;   // 2048 rows per tile

// exclusive scan of tile counts (64-bit offsets), single block, any number of tiles; 8 tiles per thread
constexpr int kScanItems = 8;
__global__ void __launch_bounds__(1024) k_scan_tiles(const int* __restrict__ tile_cnt, long long ntiles,
                                                     long long* __restrict__ tile_off) {
    pdl_prologue();
    __shared__ long long wsum[33];
    __shared__ long long carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long b = 0; b < ntiles; b += 1024 * kScanItems) {
        const long long i0 = b + (long long)threadIdx.x * kScanItems;
        long long c[kScanItems];
        long long v = 0;
#pragma unroll
        for (int t = 0; t < kScanItems; ++t) {
            c[t] = (i0 + t < ntiles) ? (long long)tile_cnt[i0 + t] : 0;
            v += c[t];
        }
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            long long w = wsum[lane];
            long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;
            if (lane == 31) wsum[32] = wi;
        }
        __syncthreads();
        long long run = carry_sh + wsum[wid] + incl - v;
#pragma unroll
        for (int t = 0; t < kScanItems; ++t) {
            if (i0 + t < ntiles) tile_off[i0 + t] = run;
            run += c[t];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = carry_sh + wsum[32];
        __syncthreads();
    }
}

}  // namespace
