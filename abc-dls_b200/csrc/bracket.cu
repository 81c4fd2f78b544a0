// bracket.cu — the two table-streaming kernels of the fast path and their small finalisers.
//
// Exact order statistics without sorting the table:
//   1. a random sample of every column (resp. of row distances) is sorted in shared memory and gives a
//      narrow *bracket* that contains the wanted order statistic with overwhelming probability;
//   2. ONE pass over the table counts the elements on either side of the brackets, appends the few
//      percent that fall inside to a candidate buffer and builds a fine histogram of the bracket;
//   3. the histogram locates the 1/1024-th of the bracket that holds the wanted rank, those few values are
//      gathered and sorted in shared memory -> the exact order statistic;
//   4. every step is verified from the counts; a miss only sets ABCDLS_S_FASTPATH_MISS (the caller reruns
//      with the 6-pass radix select of select.cu).  Results are therefore always exact or flagged.
//
// Pass A (k_bracket_pass)  : median band M and the two MAD bands L, R of all d columns in one read.
// Pass B (k_dist_pass)     : MAD-scaled Euclidean distance of every row fused with the tolerance bracket:
//                            rows with dist <= hi are emitted (row, dist) tile by tile; the N-vector `dist`
//                            is only written when the caller asks for it.
#include "common.cuh"
#include "compact.cuh"

// function attributes are set once per process (the call is not free and must not sit in the hot path)
#define ABCDLS_ONCE(h, call)             \
    do {                                 \
        static bool _done = false;       \
        if (!_done) {                    \
            ABCDLS_CUDA(h, call);        \
            _done = true;                \
        }                                \
    } while (0)

namespace {

constexpr int kNB = 1024;          // fine histogram bins per bracket
constexpr int kSample = 4096;      // sample size sorted per CTA
constexpr int kSmallCap = 8192;    // values gathered for the final in-smem sort
constexpr int kMaxD = 64;

// ---------------------------------------------------------------------------------------------------
// shared-memory bitonic sort of n = power of two doubles (n >= 8, blockDim.x * 8 >= ... any block size).
// Steps with partner distance j >= 8 work on pairs (every thread busy, two loads + two stores per pair); the last
// three steps of every merge (j = 4, 2, 1) stay inside 8 consecutive elements and run in registers.
// Three consecutive levels j = 4 jlo, 2 jlo, jlo of the merge with run length k (k >= 8 jlo) in ONE round trip through
// shared memory: a thread owns the 8 elements base + t * jlo (t = 0..7) whose indices differ exactly in the three bits
// those levels compare, and runs the three compare-exchange rounds in registers.  The sorts are shared-memory
// bandwidth bound (every level used to be a full read + write of the array), so this is ~2x.
__device__ __forceinline__ void bitonic_block8(double* s, int n, int k, int jlo) {
    const int sh = 31 - __clz(jlo);
    for (int g = threadIdx.x; g < n / 8; g += blockDim.x) {
        const int base = ((g >> sh) << (sh + 3)) | (g & (jlo - 1));
        double v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = s[base + t * jlo];
        const bool up = ((base & k) == 0);
#pragma unroll
        for (int j = 4; j > 0; j >>= 1) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if ((t & j) == 0) {
                    const double a = v[t], b = v[t | j];
                    const bool sw = (a > b) == up;
                    v[t] = sw ? b : a;
                    v[t | j] = sw ? a : b;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) s[base + t * jlo] = v[t];
    }
    __syncthreads();
}


__device__ __forceinline__ void bitonic_step(double* s, int n, int k, int j) {
    for (int p = threadIdx.x; p < n / 2; p += blockDim.x) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const double a = s[i], b = s[i | j];
        const bool up = ((i & k) == 0);
        if ((a > b) == up) {
            s[i] = b;
            s[i | j] = a;
        }
    }
    __syncthreads();
}

// sorts the first 8 elements' worth of runs (k = 2, 4, 8) entirely in registers
__device__ __forceinline__ void bitonic_head8(double* s, int n) {
    for (int blk = threadIdx.x; blk < n / 8; blk += blockDim.x) {
        double v[8];
        double* p = s + blk * 8;
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = p[t];
#pragma unroll
        for (int k = 2; k <= 8; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if ((t & j) == 0) {
                        const bool up = (((blk * 8 + t) & k) == 0);
                        const double a = v[t], b = v[t | j];
                        const bool sw = (a > b) == up;
                        v[t] = sw ? b : a;
                        v[t | j] = sw ? a : b;
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) p[t] = v[t];
    }
    __syncthreads();
}

// levels j = jtop .. 1 of one merge.  Three levels at a time (bitonic_block8) where the 8 elements of a thread are at
// least 32 apart (consecutive lanes then touch consecutive words: conflict free) and for the last three levels; pair
// steps in between (a block with a small stride would serialise on bank conflicts and lose more than it saves).
__device__ __forceinline__ void bitonic_merge_levels(double* s, int n, int k, int jtop) {
    int j = jtop;
    while (j >= 128) {            // levels j, j/2, j/4 with stride j/4 >= 32
        bitonic_block8(s, n, k, j >> 2);
        j >>= 3;
    }
    for (; j >= 8; j >>= 1) bitonic_step(s, n, k, j);
    bitonic_block8(s, n, k, 1);   // levels 4, 2, 1
}

__device__ void bitonic_sort(double* s, int n) {
    if (n < 8) {   // tiny inputs: plain network
        for (int k = 2; k <= n; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) bitonic_step(s, n, k, j);
        return;
    }
    bitonic_head8(s, n);
    for (int k = 16; k <= n; k <<= 1) bitonic_merge_levels(s, n, k, k >> 1);
}

// s is already bitonic (e.g. decreasing then increasing): one ascending merge sorts it
__device__ void bitonic_merge_up(double* s, int n) {
    bitonic_merge_levels(s, n, n << 1, n >> 1);   // k = 2n: (i & k) == 0 for every i -> ascending
}

__device__ __forceinline__ unsigned hash3(unsigned a, unsigned b, unsigned c) {
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

// random row in [0, n): 32-byte sectors of 4 consecutive doubles are drawn, so a 4096-sample costs 1024 sectors
__device__ __forceinline__ long long sample_row(unsigned seed, unsigned cta, int t, long long n) {
    unsigned h = hash3(seed, cta, (unsigned)(t >> 2));
    long long nsec = (n + 3) >> 2;
    long long sec = (long long)(((unsigned long long)h * (unsigned long long)nsec) >> 32);
    long long r = sec * 4 + (t & 3);
    return r < n ? r : n - 1;
}

// ---------------------------------------------------------------------------------------------------
// K1: per-column sample -> local estimates (m_lo, m_hi, r_lo, r_hi).  grid (G, d), 512 threads.
__global__ void __launch_bounds__(512) k_sample_cols(const double* __restrict__ ss, long long n, long long ld,
                                                     unsigned seed, double delta, double* __restrict__ est) {
    pdl_prologue();
    extern __shared__ double sm[];
    double* s = sm;            // kSample
    double* a = sm + kSample;  // kSample
    const int j = blockIdx.y;
    const double* col = ss + (size_t)j * ld;
    for (int t = threadIdx.x; t < kSample; t += blockDim.x) s[t] = col[sample_row(seed + 977u * j, blockIdx.x, t, n)];
    __syncthreads();
    bitonic_sort(s, kSample);
    const double mhat = s[kSample / 2];
    // |s - mhat| of the sorted sample falls to 0 at the median and rises again: already bitonic -> one merge
    for (int t = threadIdx.x; t < kSample; t += blockDim.x) a[t] = fabs(s[t] - mhat);
    __syncthreads();
    bitonic_merge_up(a, kSample);
    if (threadIdx.x == 0) {
        int il = (int)floor((0.5 - delta) * kSample), ih = (int)ceil((0.5 + delta) * kSample);
        il = max(0, min(kSample - 1, il));
        ih = max(0, min(kSample - 1, ih));
        double* o = est + ((size_t)blockIdx.x * gridDim.y + j) * 4;
        o[0] = s[il];
        o[1] = s[ih];
        o[2] = a[il];
        o[3] = a[ih];
    }
}

// Brackets of one column in "radial" form around a centre estimate c (NOT the exact median):
//   s = x - c,  y = |s|  (both correctly rounded; fl(x - c) is monotone in x, so every class is an interval)
//   below  : s < -r0          (strictly below the median band)
//   M band : y <= r0          (contains the two middle order statistics)
//   inner  : y <  r1          (certainly closer to the median than the MAD radius)
//   U band : r1 <= y <= r2    (contains the two middle order statistics of |x - median|)
struct ColBrk {
    double c, r0, r1, r2;
    double loM, scaleM;   // fine histogram of the M band: bin((x - loM) * scaleM), loM = c - r0
};

// K1b: average the G local estimates (already summed over ranks when world > 1) -> brackets.  One warp per column:
// lane l adds estimates l, l + 32, ... in order, then a fixed shuffle tree (deterministic).
// mean_out (optional): the averaged {m_lo, m_hi, r_lo, r_hi} of every column, for the second-stage sample of fused.cuh.
__global__ void __launch_bounds__(32) k_make_brackets(const double* __restrict__ est, int G, int d, double inv_world,
                                                      ColBrk* __restrict__ brk, double* __restrict__ mean_out = nullptr) {
    pdl_prologue();
    const int j = blockIdx.x, lane = threadIdx.x;
    if (j >= d) return;
    double m[4] = {0, 0, 0, 0};
    for (int g = lane; g < G; g += 32)
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] += est[((size_t)g * d + j) * 4 + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[q] += __shfl_down_sync(0xffffffffu, m[q], o);
        m[q] = m[q] / G * inv_world;
    }
    if (lane != 0) return;
    // m = {m_lo, m_hi, r_lo, r_hi}
    if (mean_out)
        for (int q = 0; q < 4; ++q) mean_out[j * 4 + q] = m[q];
    ColBrk B;
    B.c = 0.5 * m[0] + 0.5 * m[1];
    const double hm = fmax(fmax(m[1] - B.c, B.c - m[0]), 0.0);
    B.r0 = hm;
    B.r1 = fmax(m[2] - hm, 0.0);
    B.r2 = m[3] + hm;
    B.loM = B.c - B.r0;
    B.scaleM = (B.r0 > 0.0) ? (double)kNB / (2.0 * B.r0) : 0.0;
    brk[j] = B;
}

__device__ __forceinline__ int bin_of(double v, double lo, double scale) {
    double t = (v - lo) * scale;
    int b = (int)t;                      // v >= lo inside the band, so t >= 0
    return b < 0 ? 0 : (b >= kNB ? kNB - 1 : b);
}

// ---------------------------------------------------------------------------------------------------
// K2: pass A.  grid (gx, d), 256 threads; a CTA streams tiles of 2048 elements of ONE column.
// counts[j*4 + {0: x<b2, 1: x in M, 2: b1<x<b4, 3: x in L u R}], histM[j][kNB] over M,
// candM[j][capM] (values in M), candU[j][capU] (values in L u R); cnt[j*2 + {0,1}] = entries appended.
constexpr int kAThreads = 256;
constexpr int kAItems = 8;                    // elements per lane per warp-tile
constexpr int kAWarpTile = 32 * kAItems;      // 256 elements
constexpr int kATile = kAThreads * kAItems;   // used for grid sizing only
constexpr int kStageM = 288;                  // per-warp staging (doubles): a whole warp-tile (256) always fits
constexpr int kStageU = 384;                  //   after the flush threshold (32 / 128 staged values)

// Same-address global atomics cost ~15 ns each on B200, so candidates are staged per warp in shared memory
// and flushed with ONE atomic + a coalesced copy.  The fine histogram of the M band is built at flush time.
template <bool HIST>
__device__ __noinline__ void stage_flush(double* stage, int fill, double* __restrict__ gbuf, long long gcap,
                                         unsigned long long* gcnt, unsigned* hist, double lo, double scale) {
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0ull;
    if (lane == 0) base = atomicAdd(gcnt, (unsigned long long)fill);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int t = lane; t < fill; t += 32) {
        const double v = stage[t];
        if ((long long)base + t < gcap) gbuf[base + t] = v;
        if (HIST) atomicAdd(&hist[bin_of(v, lo, scale)], 1u);
    }
    __syncwarp();
}

// Streaming loop without block barriers.  A warp owns warp-tiles of 256 consecutive elements (8 per lane in
// registers, next tile prefetched into a second register set).  Per element: one subtract, four compares, two
// predicated counter increments; per 32 elements two ballots.  Candidate stores only happen under warp-uniform
// branches (positions from popc of the ballot), so the common path is ~8 instructions per element.
__global__ void __launch_bounds__(kAThreads, 3) k_bracket_pass(const double* __restrict__ ss, long long n, long long ld,
                                                               const ColBrk* __restrict__ brk,
                                                               unsigned long long* __restrict__ counts,
                                                               unsigned long long* __restrict__ histM,
                                                               double* __restrict__ candM, long long capM,
                                                               double* __restrict__ candU, long long capU,
                                                               unsigned long long* __restrict__ cnt) {
    pdl_prologue();
    __shared__ unsigned hist[kNB];
    __shared__ unsigned tot2[2];
    __shared__ double stM[kAThreads / 32][kStageM];
    __shared__ double stU[kAThreads / 32][kStageU];
    const int j = blockIdx.y;
    const double* __restrict__ col = ss + (size_t)j * ld;
    const ColBrk B = brk[j];
    const double c = B.c, r0 = B.r0, r1 = B.r1, r2 = B.r2, nr0 = -B.r0, loM = B.loM, scaleM = B.scaleM;
    for (int t = threadIdx.x; t < kNB; t += kAThreads) hist[t] = 0u;
    if (threadIdx.x < 2) tot2[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long warp_id = (long long)blockIdx.x * (kAThreads / 32) + wid;
    const long long warp_stride = (long long)gridDim.x * (kAThreads / 32);
    const bool vec = ((reinterpret_cast<uintptr_t>(col) & 15) == 0);
    const long long nfull = n / kAWarpTile;          // warp-tiles that need no bounds checks
    const long long ntiles = (n + kAWarpTile - 1) / kAWarpTile;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned c_lt = 0, c_in = 0;
    int fillM = 0, fillU = 0;                         // warp-uniform
    double* gM = candM + (size_t)j * capM;
    double* gU = candU + (size_t)j * capU;
    unsigned long long* cntM = &cnt[j * 2 + 0];
    unsigned long long* cntU = &cnt[j * 2 + 1];

    auto load_tile = [&](long long wt, double (&x)[kAItems]) {
        const long long t0 = wt * kAWarpTile;
        if (vec && wt < nfull) {
#pragma unroll
            for (int u = 0; u < kAItems / 2; ++u) {
                const double2 v = ld_stream2(col + t0 + u * 64 + 2 * lane);
                x[2 * u] = v.x;
                x[2 * u + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int u = 0; u < kAItems / 2; ++u) {
                const long long i = t0 + u * 64 + 2 * lane;
                x[2 * u] = i < n ? ld_stream(col + i) : nan;
                x[2 * u + 1] = i + 1 < n ? ld_stream(col + i + 1) : nan;
            }
        }
    };

#define PROCESS_TILE(x)                                                                                       \
    {                                                                                                         \
        if (fillM > kStageM - kAWarpTile) {                                                                   \
            stage_flush<true>(stM[wid], fillM, gM, capM, cntM, hist, loM, scaleM);                            \
            fillM = 0;                                                                                        \
        }                                                                                                     \
        if (fillU > kStageU - kAWarpTile) {                                                                   \
            stage_flush<false>(stU[wid], fillU, gU, capU, cntU, nullptr, 0.0, 0.0);                           \
            fillU = 0;                                                                                        \
        }                                                                                                     \
        _Pragma("unroll") for (int u = 0; u < kAItems; ++u) {                                                 \
            const double sgn = __dsub_rn(x[u], c);                                                            \
            const double y = fabs(sgn);                                                                       \
            const bool in1 = y < r1;                                                                          \
            const bool inM = y <= r0;                                                                         \
            const bool inU = !in1 && y <= r2;                                                                 \
            if (sgn < nr0) ++c_lt;                                                                            \
            if (in1) ++c_in;                                                                                  \
            const unsigned bM = __ballot_sync(0xffffffffu, inM);                                              \
            const unsigned bU = __ballot_sync(0xffffffffu, inU);                                              \
            if (bM) {                                                                                         \
                if (inM) stM[wid][fillM + __popc(bM & lt_mask)] = x[u];                                       \
                fillM += __popc(bM);                                                                          \
            }                                                                                                 \
            if (bU) {                                                                                         \
                if (inU) stU[wid][fillU + __popc(bU & lt_mask)] = x[u];                                       \
                fillU += __popc(bU);                                                                          \
            }                                                                                                 \
        }                                                                                                     \
    }

    double xa[kAItems], xb[kAItems];
    long long wt = warp_id;
    if (wt < ntiles) load_tile(wt, xa);
    while (wt < ntiles) {
        long long wnext = wt + warp_stride;
        if (wnext < ntiles) load_tile(wnext, xb);
        PROCESS_TILE(xa);
        wt = wnext;
        if (wt >= ntiles) break;
        wnext = wt + warp_stride;
        if (wnext < ntiles) load_tile(wnext, xa);
        PROCESS_TILE(xb);
        wt = wnext;
    }
#undef PROCESS_TILE
    __syncwarp();
    if (fillM) stage_flush<true>(stM[wid], fillM, gM, capM, cntM, hist, loM, scaleM);
    if (fillU) stage_flush<false>(stU[wid], fillU, gU, capU, cntU, nullptr, 0.0, 0.0);
    c_lt = warp_sum(c_lt);
    c_in = warp_sum(c_in);
    if (lane == 0) {
        atomicAdd(&tot2[0], c_lt);
        atomicAdd(&tot2[1], c_in);
    }
    __syncthreads();
    if (threadIdx.x < 2 && tot2[threadIdx.x])
        atomicAdd(&counts[j * 4 + 2 * threadIdx.x], (unsigned long long)tot2[threadIdx.x]);   // [0]=below, [2]=inner
    for (int t = threadIdx.x; t < kNB; t += kAThreads)
        if (hist[t]) atomicAdd(&histM[(size_t)j * kNB + t], (unsigned long long)hist[t]);
}

// ---------------------------------------------------------------------------------------------------
// refinement of one bracket: histogram -> the few bins that hold the wanted ranks -> gather -> sort.
struct RefJob {
    const double* src;                    // candidate values
    const unsigned long long* src_count;  // entries appended (may exceed src_cap => miss)
    long long src_cap;
    int transform;                        // 1: v = |x - center|
    int pow2k;                            // > 0: bin(v) = int(kNB * (v * lo)^(2^pow2k)) (monotone; lo = 1 / upper bound)
    int valid;                            // 0 => the bracket missed; results are NaN and MISS is flagged
    double center;
    double lo, scale;                     // bin(v) = clamp(int((v - lo) * scale))
    double vmin;                          // values below vmin are outside the bracket (not binned / gathered)
    long long kk[2];                      // 0-based ranks inside the bracket
    int nq;
    int bq[2];                            // bins holding kk[0], kk[1]          (filled by k_refine_pick)
    long long before;                     // number of bracket values in bins < bq[0]
    unsigned long long* hist;             // [kNB]
    double* small;                        // [kSmallCap]
    unsigned* small_count;
    double result[2];
};

__device__ __forceinline__ int job_bin(double v, double lo, double scale, int pow2k) {
    if (pow2k == 0) return bin_of(v, lo, scale);
    double t = __dmul_rn(v, lo);                      // v >= 0; every step is monotone in v
    for (int i = 0; i < pow2k; ++i) t = __dmul_rn(t, t);
    const int b = (int)(t * (double)kNB);
    return b < 0 ? 0 : (b >= kNB ? kNB - 1 : b);
}

// one block per job: locate the bins.  (after the histogram is complete / all-reduced)
__global__ void __launch_bounds__(256) k_refine_pick(RefJob* jobs, int* status) {
    pdl_prologue();
    RefJob& J = jobs[blockIdx.x];
    __shared__ long long wsum[9];
    __shared__ int bsel[2];
    __shared__ long long bbefore[2];
    const long long kk0 = J.kk[0], kk1 = J.kk[1];
    const bool valid = J.valid != 0;
    if (threadIdx.x < 2) { bsel[threadIdx.x] = -1; bbefore[threadIdx.x] = 0; }
    constexpr int per = kNB / 256;
    long long loc[per], s = 0;
#pragma unroll
    for (int t = 0; t < per; ++t) { loc[t] = (long long)J.hist[threadIdx.x * per + t]; s += loc[t]; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = 0;
        for (int w = 0; w < 8; ++w) { long long t = wsum[w]; wsum[w] = a; a += t; }
        wsum[8] = a;
    }
    __syncthreads();
    const long long before = wsum[wid] + incl - s;
    for (int q = 0; q < 2; ++q) {
        const long long kk = q ? kk1 : kk0;
        if (kk >= before && kk < before + s) {
            long long a = before;
            int b = 0;
            bool found = false;
#pragma unroll
            for (int t = 0; t < per; ++t) {
                if (!found) {
                    if (kk >= a + loc[t]) { a += loc[t]; b = t + 1; } else found = true;
                }
            }
            bsel[q] = threadIdx.x * per + b;
            bbefore[q] = a;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool ok = valid && bsel[0] >= 0 && bsel[1] >= 0 && kk0 >= 0 && kk1 >= kk0;
        J.bq[0] = ok ? bsel[0] : -1;
        J.bq[1] = ok ? bsel[1] : -1;
        J.before = bbefore[0];
        if (!ok) {
            J.valid = 0;
            atomicOr(status, ABCDLS_S_FASTPATH_MISS);
        }
        *J.small_count = 0u;
    }
}

// grid (gx, njobs): append the values that fall in bins [bq0, bq1] to J.small
__global__ void __launch_bounds__(256) k_refine_gather(RefJob* jobs, long long n_max) {
    pdl_prologue();
    RefJob& J = jobs[blockIdx.y];
    if (!J.valid) return;
    long long n = (long long)*J.src_count;
    if (n > J.src_cap) n = J.src_cap;   // overflow was flagged by the plan kernel
    if (n > n_max) n = n_max;
    const int b0 = J.bq[0], b1 = J.bq[1];
    const double lo = J.lo, scale = J.scale, c = J.center, vmin = J.vmin;
    const bool tr = J.transform != 0;
    const int pow2k = J.pow2k;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // four independent streaming loads in flight per thread (the candidate lists are read once, from HBM)
    const double* __restrict__ src = J.src;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            xv[u] = i < n ? ld_stream(src + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * stride >= n) continue;
            const double v = tr ? fabs(__dsub_rn(xv[u], c)) : xv[u];
            if (v >= vmin) {
                const int b = job_bin(v, lo, scale, pow2k);
                if (b >= b0 && b <= b1) {
                    unsigned p = atomicAdd(J.small_count, 1u);
                    if (p < (unsigned)kSmallCap) J.small[p] = v;
                }
            }
        }
    }
}

// one block (1024 threads) per job: sort the gathered values, read the ranks.
// one block (1024 threads) per job: sort the gathered values, read the ranks.
// gsmall == NULL: the job's own buffer.  Otherwise the all-gathered exchange block of this phase,
// gsmall[world][nj][kSmallCap] / gcnt[world][nj], job = blockIdx.x within the phase.
__global__ void __launch_bounds__(1024) k_refine_final(RefJob* jobs, const double* __restrict__ gsmall,
                                                       const unsigned* __restrict__ gcnt, int world, int nj,
                                                       int* status) {
    pdl_prologue();
    extern __shared__ double sm[];
    __shared__ unsigned off[65];
    RefJob& J = jobs[blockIdx.x];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    bool ok = J.valid != 0 && world <= 64;
    if (threadIdx.x == 0) {
        unsigned a = 0;
        for (int r = 0; r < world; ++r) {
            off[r] = a;
            const unsigned c = gsmall ? gcnt[r * nj + blockIdx.x] : *J.small_count;
            a += c;
        }
        off[world] = a;
    }
    __syncthreads();
    const unsigned cnt = off[world];
    ok = ok && cnt <= (unsigned)kSmallCap && cnt > 0;
    const long long r0 = J.kk[0] - J.before, r1 = J.kk[1] - J.before;
    ok = ok && r0 >= 0 && r1 < (long long)cnt;
    if (!ok) {
        if (threadIdx.x == 0) {
            J.result[0] = J.result[1] = nan;
            atomicOr(status, ABCDLS_S_FASTPATH_MISS);
        }
        return;
    }
    int np2 = 32;
    while (np2 < (int)cnt) np2 <<= 1;
    for (int t = threadIdx.x; t < np2; t += blockDim.x) sm[t] = inf;
    __syncthreads();
    for (int r = 0; r < world; ++r) {
        const unsigned c = off[r + 1] - off[r];
        const double* src = gsmall ? gsmall + ((size_t)r * nj + blockIdx.x) * kSmallCap : J.small;
        for (int t = threadIdx.x; t < (int)c; t += blockDim.x) sm[off[r] + t] = src[t];
    }
    __syncthreads();
    bitonic_sort(sm, np2);
    if (threadIdx.x == 0) {
        J.result[0] = sm[r0];
        J.result[1] = sm[r1];
    }
}

__global__ void k_copy_u64(unsigned long long* dst, const unsigned long long* src, int n) {
    pdl_prologue();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------
// plans (1 thread per column).  k0/k1: global 0-based ranks of the two middle order statistics.
__global__ void k_plan_median(RefJob* jobs, int d, const ColBrk* __restrict__ brk,
                              const unsigned long long* __restrict__ counts, unsigned long long* histM,
                              const double* candM, long long capM, const unsigned long long* cnt,
                              const unsigned long long* gcnt, double* small, unsigned* small_count, long long k0,
                              long long k1, int* status) {
    pdl_prologue();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    RefJob J;
    // counts / gcnt are summed over ranks; cnt is what THIS rank appended (every band value is appended)
    const long long c_lt = (long long)counts[j * 4 + 0], c_m = (long long)gcnt[j * 2 + 0];
    J.src = candM + (size_t)j * capM;
    J.src_count = cnt + j * 2 + 0;
    J.src_cap = capM;
    J.transform = 0;
    J.pow2k = 0;
    J.center = 0.0;
    J.lo = brk[j].loM;
    J.scale = brk[j].scaleM;
    J.vmin = -__longlong_as_double(0x7ff0000000000000ll);
    J.kk[0] = k0 - c_lt;
    J.kk[1] = k1 - c_lt;
    J.nq = 2;
    J.bq[0] = J.bq[1] = -1;
    J.before = 0;
    J.hist = histM + (size_t)j * kNB;
    J.small = small + (size_t)j * kSmallCap;
    J.small_count = small_count + j;
    J.result[0] = J.result[1] = 0.0;
    // the band must contain both ranks and every band value must have been stored
    J.valid = (J.kk[0] >= 0 && J.kk[1] < c_m && (long long)cnt[j * 2 + 0] <= J.src_cap) ? 1 : 0;
    if (!J.valid) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    jobs[j] = J;
}

// after the medians are exact: med[j] = 0.5*r0 + 0.5*r1; prepare the |x - med| refinement over L u R
__global__ void k_plan_mad(RefJob* jobs, int d, const ColBrk* __restrict__ brk,
                           const unsigned long long* __restrict__ counts, unsigned long long* histU,
                           const double* candU, long long capU, const unsigned long long* cnt,
                           const unsigned long long* gcnt, double* small, unsigned* small_count, long long k0,
                           long long k1, double* __restrict__ med, double* __restrict__ guard, int* status) {
    pdl_prologue();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const RefJob& Jm = jobs[j];
    const double m = __dadd_rn(__dmul_rn(0.5, Jm.result[0]), __dmul_rn(0.5, Jm.result[1]));
    med[j] = m;
    const ColBrk B = brk[j];
    RefJob J;
    const long long c_in = (long long)counts[j * 4 + 2], c_u = (long long)gcnt[j * 2 + 1];
    J.src = candU + (size_t)j * capU;
    J.src_count = cnt + j * 2 + 1;
    J.src_cap = capU;
    J.transform = 1;
    J.pow2k = 0;
    J.center = m;
    // x in U has fl|x - c| in [r1, r2]; with e = |c - m| its |x - m| lies in [r1 - e, r2 + e] up to rounding
    const double e = fabs(__dsub_rn(B.c, m));
    const double up = 1.0 + 0x1p-48, dn = 1.0 - 0x1p-48;
    const double alo = fmax(0.0, (B.r1 - e) * dn);
    const double ahi = (B.r2 + e) * up;
    J.lo = alo;
    J.scale = (ahi > alo) ? (double)kNB / (ahi - alo) : 0.0;
    J.vmin = -__longlong_as_double(0x7ff0000000000000ll);
    J.kk[0] = k0 - c_in;
    J.kk[1] = k1 - c_in;
    J.nq = 2;
    J.bq[0] = J.bq[1] = -1;
    J.before = 0;
    J.hist = histU + (size_t)j * kNB;
    J.small = small + (size_t)(d + j) * kSmallCap;
    J.small_count = small_count + d + j;
    J.result[0] = J.result[1] = 0.0;
    J.valid = (Jm.valid && J.kk[0] >= 0 && J.kk[1] < c_u && (long long)cnt[j * 2 + 1] <= J.src_cap) ? 1 : 0;
    if (!J.valid) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    // guards checked once r* is known (triangle inequality + rounding slack):
    //   inner rows (fl|x-c| < r1)  have |x - m| <= (r1 + e) * up   -> must be <= the smaller rank value
    //   outer rows (fl|x-c| > r2)  have |x - m| >= (r2 - e) * dn   -> must be >= the larger rank value
    guard[j * 2 + 0] = (B.r1 > 0.0) ? (B.r1 + e) * up : -1.0;
    guard[j * 2 + 1] = (B.r2 - e) * dn;
    jobs[d + j] = J;
}

// histogram of v = |x - center| (or x) of the candidates of every job.  grid (gx, njobs)
__global__ void __launch_bounds__(256) k_refine_hist(RefJob* jobs, long long n_max) {
    pdl_prologue();
    __shared__ unsigned hist[kNB];
    RefJob& J = jobs[blockIdx.y];
    if (!J.valid) return;
    for (int t = threadIdx.x; t < kNB; t += blockDim.x) hist[t] = 0u;
    __syncthreads();
    long long n = (long long)*J.src_count;
    if (n > J.src_cap) n = J.src_cap;
    if (n > n_max) n = n_max;
    const double lo = J.lo, scale = J.scale, c = J.center, vmin = J.vmin;
    const bool tr = J.transform != 0;
    const int pow2k = J.pow2k;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double* __restrict__ src = J.src;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            xv[u] = i < n ? ld_stream(src + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * stride >= n) continue;
            const double v = tr ? fabs(__dsub_rn(xv[u], c)) : xv[u];
            if (v >= vmin) atomicAdd(&hist[job_bin(v, lo, scale, pow2k)], 1u);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kNB; t += blockDim.x)
        if (hist[t]) atomicAdd(&J.hist[t], (unsigned long long)hist[t]);
}

__global__ void k_finish_mad(const RefJob* jobs, int d, const double* __restrict__ guard, double* __restrict__ mad,
                             int* status) {
    pdl_prologue();
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    const RefJob& J = jobs[d + j];
    const double r0 = J.result[0], r1 = J.result[1];
    // r0 <= r1 are the two middle order statistics of |x - med|
    const bool ok = J.valid && (guard[j * 2 + 0] <= r0) && (guard[j * 2 + 1] >= r1);
    if (!ok) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    mad[j] = __dmul_rn(1.4826, __dadd_rn(__dmul_rn(0.5, r0), __dmul_rn(0.5, r1)));
}

// ---------------------------------------------------------------------------------------------------
// distance: shared scaling constants
struct ColScale {
    double2 dr[kMaxD];   // (divisor, RN(1/divisor)); rcp = 0 marks a column whose divisor is outside [1e-90, 1e90]
    double tgt[kMaxD];   // target / divisor
};

__device__ __forceinline__ void load_scale(ColScale& S, const double* mad, const double* target, int d) {
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        double m = mad[j];
        double dv = (m == 0.0) ? 1.0 : m;
        double a = fabs(dv);
        bool sane = a > 1e-90 && a < 1e90;
        S.dr[j] = make_double2(dv, sane ? 1.0 / dv : 0.0);
        S.tgt[j] = (m == 0.0) ? target[j] : target[j] / dv;
    }
}

// x / div, correctly rounded.  Markstein's sequence is exact while no intermediate leaves the normal range:
// |x| in (1e-200, 1e200) with |div| in (1e-90, 1e90) keeps q, the remainders and their products normal; x == 0 is
// exact too.  Everything else takes the IEEE division.
__device__ __forceinline__ double scaled_value(double x, const double2 dr) {
    const double ax = fabs(x);
    if ((ax > 1e-200 && ax < 1e200 && dr.y != 0.0) || x == 0.0) return div_const(x, dr.x, dr.y);
    return x / dr.x;
}

// K6: distances of a random sample of rows -> local estimates (lo, hi) of the nacc-th smallest distance.
// grid G, 512 threads; est[g*2 + {0,1}]
__global__ void __launch_bounds__(512) k_sample_dist(const double* __restrict__ ss, long long n, int d, long long ld,
                                                     const double* __restrict__ mad, const double* __restrict__ target,
                                                     unsigned seed, double p, double delta, double* __restrict__ est) {
    pdl_prologue();
    extern __shared__ double sm[];
    __shared__ ColScale S;
    load_scale(S, mad, target, d);
    __syncthreads();
    for (int t = threadIdx.x; t < kSample; t += blockDim.x) {
        long long r = sample_row(seed, blockIdx.x, t, n);
        double s0 = 0.0;
        for (int j = 0; j < d; ++j) {
            double a = __dsub_rn(scaled_value(ss[(size_t)j * ld + r], S.dr[j]), S.tgt[j]);
            s0 = __dadd_rn(s0, __dmul_rn(a, a));
        }
        sm[t] = __dsqrt_rn(s0);
    }
    __syncthreads();
    bitonic_sort(sm, kSample);
    if (threadIdx.x == 0) {
        int il = (int)floor((p - delta) * kSample) - 1, ih = (int)ceil((p + delta) * kSample) + 1;
        il = max(0, min(kSample - 1, il));
        ih = max(0, min(kSample - 1, ih));
        est[blockIdx.x * 2 + 0] = (il == 0) ? 0.0 : sm[il];
        est[blockIdx.x * 2 + 1] = sm[ih];
    }
}

struct DistBrk {
    double lo, hi, scale;
};

__global__ void k_make_dist_bracket(const double* __restrict__ est, int G, double inv_world, DistBrk* brk) {
    pdl_prologue();
    if (threadIdx.x || blockIdx.x) return;
    double lo = 0, hi = 0;
    for (int g = 0; g < G; ++g) { lo += est[g * 2]; hi += est[g * 2 + 1]; }
    lo = lo / G * inv_world;
    hi = hi / G * inv_world;
    brk->lo = lo;
    brk->hi = hi;
    brk->scale = (hi > lo) ? (double)kNB / (hi - lo) : 0.0;
}

// K7: pass B.  Every warp owns one contiguous range of rows and walks it in warp-tiles of 128 rows (lane l holds
// rows {2l, 2l+1} of each 64-row half: two 16-byte loads per column).  Candidate rows (dist <= hi) are staged
// per warp in shared memory in row order and flushed with one atomic + coalesced copies; each flush leaves a
// record (offset, count) so that k_order_cands can lay all chunks out in global row order.  No block barriers.
constexpr int kBThreads = 256;
constexpr int kBGrid = 148 * 8;                 // CTAs (fixed: the record directory is sized from it)
constexpr int kBWarps = kBGrid * (kBThreads / 32);
constexpr int kBHalf = 64;
constexpr int kBWarpTile = 128;
constexpr int kStageB = 96;                     // staged candidates per warp

inline long long rows_per_warp(long long n) {
    long long r = (n + kBWarps - 1) / kBWarps;
    return (r + kBWarpTile - 1) / kBWarpTile * kBWarpTile;
}
inline long long recs_per_warp(long long n) { return 2 * (rows_per_warp(n) / kBWarpTile) + 2; }

template <bool VEC2, bool WRITE_DIST>
__global__ void __launch_bounds__(kBThreads) k_dist_pass(const double* __restrict__ ss, long long n, int d, long long ld,
                                                         const double* __restrict__ mad, const double* __restrict__ target,
                                                         const DistBrk* __restrict__ brk, double* __restrict__ dist,
                                                         long long* __restrict__ cand_row, double* __restrict__ cand_dist,
                                                         long long cap, unsigned long long* __restrict__ cand_cnt,
                                                         unsigned long long* __restrict__ n_lt,
                                                         unsigned long long* __restrict__ hist_g,
                                                         long long row_lo, int prod0, long long rpw, long long rcap,
                                                         long long* __restrict__ rec_off,
                                                         int* __restrict__ rec_cnt, int* __restrict__ wtot,
                                                         int* __restrict__ nrec, int* status) {
    pdl_prologue();
    __shared__ ColScale S;
    __shared__ unsigned hist[kNB];
    __shared__ long long st_row[kBThreads / 32][kStageB];
    __shared__ double st_dist[kBThreads / 32][kStageB];
    load_scale(S, mad, target, d);
    for (int t = threadIdx.x; t < kNB; t += kBThreads) hist[t] = 0u;
    __syncthreads();
    const double lo = brk->lo, hi = brk->hi, scale = brk->scale;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long gw0 = (long long)blockIdx.x * (kBThreads / 32) + wid;
    const long long gw = prod0 + gw0;   // slot in the record directory
    const long long r_begin = min(n, row_lo + gw0 * rpw), r_end = min(n, r_begin + rpw);
    bool nonfinite = false;
    unsigned c_lt = 0;
    int fill = 0, seq = 0, total = 0;

    auto flush = [&]() {
        unsigned long long base = 0ull;
        if (lane == 0) {
            base = atomicAdd(cand_cnt, (unsigned long long)fill);
            if (seq < rcap) {
                rec_off[gw * rcap + seq] = (long long)base;
                rec_cnt[gw * rcap + seq] = fill;
            }
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int t = lane; t < fill; t += 32)
            if ((long long)base + t < cap) {
                cand_row[base + t] = st_row[wid][t];
                cand_dist[base + t] = st_dist[wid][t];
            }
        ++seq;
        fill = 0;
        __syncwarp();
    };

    for (long long t0 = r_begin; t0 < r_end; t0 += kBWarpTile) {
        const long long i0 = t0 + 2 * lane, i1 = i0 + kBHalf;
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        const bool full = VEC2 && (t0 + kBWarpTile <= n);
        if (full) {
            const double* col = ss + i0;
#pragma unroll 4
            for (int j = 0; j < d; ++j) {
                const double2 xa = ld_stream2(col);
                const double2 xb = ld_stream2(col + kBHalf);
                col += ld;
                const double tj = S.tgt[j];
                const double2 dr = S.dr[j];
                const double a0 = __dsub_rn(scaled_value(xa.x, dr), tj);
                const double a1 = __dsub_rn(scaled_value(xa.y, dr), tj);
                const double a2 = __dsub_rn(scaled_value(xb.x, dr), tj);
                const double a3 = __dsub_rn(scaled_value(xb.y, dr), tj);
                s[0] = __dadd_rn(s[0], __dmul_rn(a0, a0));
                s[1] = __dadd_rn(s[1], __dmul_rn(a1, a1));
                s[2] = __dadd_rn(s[2], __dmul_rn(a2, a2));
                s[3] = __dadd_rn(s[3], __dmul_rn(a3, a3));
            }
        } else {
            const long long rows[4] = {i0, i0 + 1, i1, i1 + 1};
            for (int j = 0; j < d; ++j) {
                const double tj = S.tgt[j];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rows[q] < n) {
                        const double a = __dsub_rn(scaled_value(ss[(size_t)j * ld + rows[q]], S.dr[j]), tj);
                        s[q] = __dadd_rn(s[q], __dmul_rn(a, a));
                    }
            }
        }
        const long long rows[4] = {i0, i0 + 1, i1, i1 + 1};
        double dd[4];
        bool c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dd[q] = __dsqrt_rn(s[q]);
            const bool v = rows[q] < n;
            nonfinite |= v && !isfinite(dd[q]);
            c[q] = v && dd[q] <= hi;
            c_lt += (v && dd[q] < lo) ? 1u : 0u;
        }
        if (WRITE_DIST) {
            if (full) {
                *reinterpret_cast<double2*>(dist + i0) = make_double2(dd[0], dd[1]);
                *reinterpret_cast<double2*>(dist + i1) = make_double2(dd[2], dd[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rows[q] < n) dist[rows[q]] = dd[q];
            }
        }
        const int packed = ((c[0] ? 1 : 0) + (c[1] ? 1 : 0)) | (((c[2] ? 1 : 0) + (c[3] ? 1 : 0)) << 8);
        if (__any_sync(0xffffffffu, packed != 0)) {
            int incl = packed;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int tot = __shfl_sync(0xffffffffu, incl, 31);
            const int tot0 = tot & 0xff, tot1 = tot >> 8;
            const int ex = incl - packed;
            // the two halves are appended one after the other (row order); each is at most 64 <= kStageB
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int th = half ? tot1 : tot0;
                if (th == 0) continue;
                if (fill + th > kStageB) flush();
                int p = fill + (half ? (ex >> 8) : (ex & 0xff));
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int qq = 2 * half + q;
                    if (c[qq]) {
                        st_row[wid][p] = rows[qq];
                        st_dist[wid][p] = dd[qq];
                        ++p;
                        if (dd[qq] >= lo) atomicAdd(&hist[bin_of(dd[qq], lo, scale)], 1u);
                    }
                }
                fill += th;
                total += th;
                __syncwarp();
            }
        }
    }
    if (fill > 0) flush();
    if (lane == 0) {
        wtot[gw] = total;
        nrec[gw] = seq < rcap ? seq : (int)rcap;
        if (seq > rcap) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    }
    c_lt = warp_sum(c_lt);
    if (lane == 0 && c_lt) atomicAdd(n_lt, (unsigned long long)c_lt);
    if (__any_sync(0xffffffffu, nonfinite) && lane == 0) atomicOr(status, ABCDLS_S_NONFINITE);
    __syncthreads();
    for (int t = threadIdx.x; t < kNB; t += kBThreads)
        if (hist[t]) atomicAdd(&hist_g[t], (unsigned long long)hist[t]);
}

// K7-TMA: the same pass with the table streamed by the TMA engine.  One producer warp issues 1-D bulk copies
// (cp.async.bulk, 2 KB per column per tile) into a 3-stage shared-memory ring guarded by mbarriers; 8 consumer
// warps own one row each of the 256-row tile and never wait on HBM latency themselves, so the bytes in flight
// (up to 96 KB per CTA, 2 CTAs per SM) no longer depend on registers.  Tiles are handed to CTAs in contiguous
// ranges; per tile the 8 consumer warps meet once at a named barrier to append their candidates in row order to
// a CTA-level stage, which is flushed with one atomic + coalesced copies and leaves a record for k_order_cands.
constexpr int kTRows = 256;                  // rows per tile = consumer threads
constexpr int kTCols = 16;                   // columns per ring stage
constexpr int kTStages = 3;
constexpr int kTConsumers = 256;
constexpr int kTThreads = kTConsumers + 32;  // + producer warp
constexpr int kTGrid = 148 * 2;
constexpr int kTStageCta = 512;              // staged candidates per CTA (a tile adds at most 256)

struct TmaSmem {
    double ring[kTStages][kTCols][kTRows];   // 96 KB
    long long st_row[kTStageCta];
    double st_dist[kTStageCta];
    unsigned hist[kNB];
    ColScale S;
    unsigned long long full[kTStages], empty[kTStages];
    int warp_cnt[2][8];
    unsigned long long base_sh;
};

static_assert(sizeof(TmaSmem) <= 113 * 1024, "two CTAs of the TMA distance kernel must fit one SM");

inline long long tma_tiles_per_cta(long long ntiles) { return (ntiles + kTGrid - 1) / kTGrid; }
inline long long tma_recs_per_cta(long long ntiles) { return tma_tiles_per_cta(ntiles) / 2 + 3; }

template <bool WRITE_DIST>
__global__ void __launch_bounds__(kTThreads, 2) k_dist_pass_tma(const double* __restrict__ ss, long long ntiles, int d,
                                                                long long ld, const double* __restrict__ mad,
                                                                const double* __restrict__ target,
                                                                const DistBrk* __restrict__ brk, double* __restrict__ dist,
                                                                long long* __restrict__ cand_row,
                                                                double* __restrict__ cand_dist, long long cap,
                                                                unsigned long long* __restrict__ cand_cnt,
                                                                unsigned long long* __restrict__ n_lt,
                                                                unsigned long long* __restrict__ hist_g, long long tpc,
                                                                long long rcap, long long* __restrict__ rec_off,
                                                                int* __restrict__ rec_cnt, int* __restrict__ wtot,
                                                                int* __restrict__ nrec, int* status) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TmaSmem& sm = *reinterpret_cast<TmaSmem*>(smem_raw);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_scale(sm.S, mad, target, d);
    for (int t = threadIdx.x; t < kNB; t += kTThreads) sm.hist[t] = 0u;
    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < kTStages; ++s2) {
            mbar_init(&sm.full[s2], 1);
            mbar_init(&sm.empty[s2], kTConsumers / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const long long t_begin = (long long)blockIdx.x * tpc, t_end = min(ntiles, t_begin + tpc);
    const int nchunk = (d + kTCols - 1) / kTCols;

    if (wid == kTConsumers / 32) {
        // ===== producer warp: one elected lane feeds the ring =====
        if (lane == 0) {
            long long it = 0;
            for (long long tile = t_begin; tile < t_end; ++tile) {
                const double* src0 = ss + tile * kTRows;
                for (int ch = 0; ch < nchunk; ++ch, ++it) {
                    const int st = (int)(it % kTStages);
                    const unsigned ph = (unsigned)((it / kTStages) & 1);
                    mbar_wait(&sm.empty[st], ph ^ 1u);
                    const int j0 = ch * kTCols, nc = min(kTCols, d - j0);
                    mbar_expect_tx(&sm.full[st], (unsigned)(nc * kTRows * 8));
                    for (int c = 0; c < nc; ++c)
                        bulk_g2s(&sm.ring[st][c][0], src0 + (size_t)(j0 + c) * ld, kTRows * 8, &sm.full[st]);
                }
            }
        }
        return;
    }

    // ===== consumers: thread t owns row t of every tile =====
    const double lo = brk->lo, hi = brk->hi, scale = brk->scale;
    const int row_in_tile = threadIdx.x;   // 0..255
    bool nonfinite = false;
    unsigned c_lt = 0;
    int fill = 0, seq = 0, total = 0;
    long long it = 0;

    auto flush = [&]() {   // all 256 consumers; every earlier append is visible (it happened before this tile's barrier)
        if (threadIdx.x == 0) {
            const unsigned long long b = atomicAdd(cand_cnt, (unsigned long long)fill);
            sm.base_sh = b;
            if (seq < rcap) {
                rec_off[(long long)blockIdx.x * rcap + seq] = (long long)b;
                rec_cnt[(long long)blockIdx.x * rcap + seq] = fill;
            }
        }
        named_bar_sync(1, kTConsumers);
        const long long base = (long long)sm.base_sh;
        for (int t = threadIdx.x; t < fill; t += kTConsumers)
            if (base + t < cap) {
                cand_row[base + t] = sm.st_row[t];
                cand_dist[base + t] = sm.st_dist[t];
            }
        named_bar_sync(1, kTConsumers);
        ++seq;
        fill = 0;
    };

    for (long long tile = t_begin; tile < t_end; ++tile) {
        double s0 = 0.0;
        for (int ch = 0; ch < nchunk; ++ch, ++it) {
            const int st = (int)(it % kTStages);
            const unsigned ph = (unsigned)((it / kTStages) & 1);
            mbar_wait(&sm.full[st], ph);
            const int j0 = ch * kTCols, nc = min(kTCols, d - j0);
            if (nc == kTCols) {
#pragma unroll
                for (int c = 0; c < kTCols; ++c) {
                    const double x = sm.ring[st][c][row_in_tile];
                    const double a = __dsub_rn(scaled_value(x, sm.S.dr[j0 + c]), sm.S.tgt[j0 + c]);
                    s0 = __dadd_rn(s0, __dmul_rn(a, a));
                }
            } else {
                for (int c = 0; c < nc; ++c) {
                    const double x = sm.ring[st][c][row_in_tile];
                    const double a = __dsub_rn(scaled_value(x, sm.S.dr[j0 + c]), sm.S.tgt[j0 + c]);
                    s0 = __dadd_rn(s0, __dmul_rn(a, a));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[st]);   // this warp is done with the stage
        }
        const long long row = tile * kTRows + row_in_tile;
        const double dd = __dsqrt_rn(s0);
        nonfinite |= !isfinite(dd);
        if (WRITE_DIST) dist[row] = dd;
        const bool cnd = dd <= hi;
        c_lt += (dd < lo) ? 1u : 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, cnd);
        const int par = (int)(tile & 1);
        if (lane == 0) sm.warp_cnt[par][wid] = __popc(bal);
        named_bar_sync(1, kTConsumers);
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int t = sm.warp_cnt[par][w];
            pre += (w < wid) ? t : 0;
            tot += t;
        }
        if (tot) {
            if (fill + tot > kTStageCta) flush();
            if (cnd) {
                const int p = fill + pre + __popc(bal & ((1u << lane) - 1u));
                sm.st_row[p] = row;
                sm.st_dist[p] = dd;
                if (dd >= lo) atomicAdd(&sm.hist[bin_of(dd, lo, scale)], 1u);
            }
            fill += tot;
            total += tot;
        }
    }
    named_bar_sync(1, kTConsumers);   // appends of the last tile are visible
    if (fill > 0) flush();
    if (threadIdx.x == 0) {
        wtot[blockIdx.x] = total;
        nrec[blockIdx.x] = seq < rcap ? seq : (int)rcap;
        if (seq > rcap) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    }
    c_lt = warp_sum(c_lt);
    if (lane == 0 && c_lt) atomicAdd(n_lt, (unsigned long long)c_lt);
    if (__any_sync(0xffffffffu, nonfinite) && lane == 0) atomicOr(status, ABCDLS_S_NONFINITE);
    named_bar_sync(1, kTConsumers);
    for (int t = threadIdx.x; t < kNB; t += kTConsumers)
        if (sm.hist[t]) atomicAdd(&hist_g[t], (unsigned long long)sm.hist[t]);
}

// K8: lay the flushed chunks out in global row order.  One warp per producer warp; wpos = exclusive scan of wtot.
__global__ void __launch_bounds__(256) k_order_cands(const long long* __restrict__ rec_off, const int* __restrict__ rec_cnt,
                                                     const int* __restrict__ nrec, const long long* __restrict__ wpos,
                                                     int nprod, long long rcap, const long long* __restrict__ row_in,
                                                     const double* __restrict__ dist_in, long long cap,
                                                     long long row_offset, long long* __restrict__ row_out,
                                                     double* __restrict__ dist_out) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (gw >= nprod) return;
    long long dst = wpos[gw];
    const int nr = nrec[gw];
    for (int r = 0; r < nr; ++r) {
        const long long src = rec_off[gw * rcap + r];
        const int c = rec_cnt[gw * rcap + r];
        for (int e = lane; e < c; e += 32)
            if (src + e < cap && dst + e < cap) {
                row_out[dst + e] = row_in[src + e] + row_offset;
                dist_out[dst + e] = dist_in[src + e];
            }
        dst += c;
    }
}

__global__ void k_plan_ds(RefJob* job, const DistBrk* __restrict__ brk, const unsigned long long* __restrict__ n_lt,
                          const unsigned long long* __restrict__ cand_cnt,
                          const unsigned long long* __restrict__ gcand_cnt, unsigned long long* hist,
                          const double* cand_dist, long long cap, double* small, unsigned* small_count, long long k,
                          int* status) {
    pdl_prologue();
    if (threadIdx.x || blockIdx.x) return;
    RefJob J;
    const long long c_lt = (long long)*n_lt, c_all = (long long)*gcand_cnt;   // summed over ranks
    J.src = cand_dist;
    J.src_count = cand_cnt;
    J.src_cap = cap;
    J.transform = 0;
    J.pow2k = 0;
    J.center = 0.0;
    J.lo = brk->lo;
    J.scale = brk->scale;
    J.vmin = brk->lo;
    J.kk[0] = J.kk[1] = k - c_lt;
    J.nq = 1;
    J.bq[0] = J.bq[1] = -1;
    J.before = 0;
    J.hist = hist;
    J.small = small;
    J.small_count = small_count;
    J.result[0] = J.result[1] = 0.0;
    J.valid = (k >= c_lt && k < c_all && (long long)*cand_cnt <= cap) ? 1 : 0;
    if (!J.valid) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    *job = J;
}

__global__ void k_copy_result(const RefJob* job, double* out) {
    pdl_prologue();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = job->result[0];
}

// ---------------------------------------------------------------------------------------------------
// workspace carving
struct MadWs {
    double* est;                    // [G*d*4]
    ColBrk* brk;                    // [d]
    unsigned long long* cnt;        // [d*2]   entries appended by THIS rank   | zeroed together
    unsigned long long* counts;     // [d*4]   } reduce block A (summed        |
    unsigned long long* gcnt;       // [d*2]   }  over ranks after the pass)   |
    unsigned long long* histM;      // [d*kNB] }                               |
    unsigned long long* histU;      // [d*kNB] reduce block B                  |
    unsigned* small_count;          // [2d]                                    |
    size_t zero_bytes, redA_count, redB_count;
    RefJob* jobs;                   // [2d]
    double* guard;                  // [2d]
    double* small;                  // [2d*kSmallCap]
    double* candM;                  // [d*capM]
    double* candU;                  // [d*capU]
    size_t total;
};

inline MadWs carve_mad(void* ws, int d, int G, long long capM, long long capU) {
    MadWs w;
    char* p = reinterpret_cast<char*>(ws);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t a = o; o += align_up(bytes, 256); return p ? p + a : (char*)nullptr; };
    w.est = (double*)take((size_t)G * d * 4 * 8);
    w.brk = (ColBrk*)take((size_t)d * sizeof(ColBrk));
    size_t z0 = o;
    w.cnt = (unsigned long long*)take((size_t)d * 2 * 8);
    size_t a0 = o;
    w.counts = (unsigned long long*)take((size_t)d * 4 * 8);
    w.gcnt = (unsigned long long*)take((size_t)d * 2 * 8);
    w.histM = (unsigned long long*)take((size_t)d * kNB * 8);
    size_t b0 = o;
    w.histU = (unsigned long long*)take((size_t)d * kNB * 8);
    size_t b1 = o;
    w.small_count = (unsigned*)take((size_t)2 * d * 4);
    w.zero_bytes = o - z0;
    w.redA_count = (b0 - a0) / 8;
    w.redB_count = (b1 - b0) / 8;
    w.jobs = (RefJob*)take((size_t)2 * d * sizeof(RefJob));
    w.guard = (double*)take((size_t)2 * d * 8);
    w.small = (double*)take((size_t)2 * d * kSmallCap * 8);
    w.candM = (double*)take((size_t)d * capM * 8);
    w.candU = (double*)take((size_t)d * capU * 8);
    w.total = o;
    return w;
}

struct DistWs {
    double* est;                   // [G*2]
    DistBrk* brk;
    unsigned long long* cand_cnt;  // candidates appended by THIS rank  | zeroed together
    unsigned long long* gcand_cnt; // } reduce block (summed over ranks) |
    unsigned long long* n_lt;      // }                                  |
    unsigned long long* hist;      // } [kNB]                            |
    unsigned* small_count;         //                                    |
    size_t zero_bytes, red_count;
    RefJob* job;
    double* small;                 // [kSmallCap]
    long long* rec_off;            // [kBWarps * rcap]
    int* rec_cnt;                  // [kBWarps * rcap]
    int* wtot;                     // [kBWarps]
    int* nrec;                     // [kBWarps]
    long long* wpos;               // [kBWarps]
    long long* row_tmp;            // [cap]
    double* dist_tmp;              // [cap]
    size_t total;
};

inline DistWs carve_dist(void* ws, int G, long long n, long long cap) {
    DistWs w;
    char* p = reinterpret_cast<char*>(ws);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t a = o; o += align_up(bytes, 256); return p ? p + a : (char*)nullptr; };
    const long long dir_ldg = (long long)kBWarps * recs_per_warp(n);
    const long long dir_tma = (long long)(kTGrid + 8) * tma_recs_per_cta(n / kTRows + 1);
    const long long dir_entries = dir_ldg > dir_tma ? dir_ldg : dir_tma;
    w.est = (double*)take((size_t)G * 2 * 8);
    w.brk = (DistBrk*)take(sizeof(DistBrk));
    size_t z0 = o;
    w.cand_cnt = (unsigned long long*)take(8);
    size_t r0 = o;
    w.gcand_cnt = (unsigned long long*)take(8);
    w.n_lt = (unsigned long long*)take(8);
    w.hist = (unsigned long long*)take((size_t)kNB * 8);
    size_t r1 = o;
    w.small_count = (unsigned*)take(4);
    w.zero_bytes = o - z0;
    w.red_count = (r1 - r0) / 8;
    w.job = (RefJob*)take(sizeof(RefJob));
    w.small = (double*)take((size_t)kSmallCap * 8);
    w.rec_off = (long long*)take((size_t)dir_entries * 8);
    w.rec_cnt = (int*)take((size_t)dir_entries * 4);
    w.wtot = (int*)take((size_t)kBWarps * 4);
    w.nrec = (int*)take((size_t)kBWarps * 4);
    w.wpos = (long long*)take((size_t)kBWarps * 8);
    w.row_tmp = (long long*)take((size_t)cap * 8);
    w.dist_tmp = (double*)take((size_t)cap * 8);
    w.total = o;
    return w;
}

inline int sample_ctas(long long n) {
    // total sample ~ n/64, between 8 and 64 CTAs of 4096 per column
    long long g = n / 64 / kSample;
    if (g < 8) g = 8;
    if (g > 64) g = 64;
    return (int)g;
}

// The estimates of all ranks are averaged, so each of `world` ranks only has to contribute 1/world of the sample:
// the sampling kernels (a fixed cost that does not shrink with the shard) get `world` times cheaper.
inline int sample_ctas_w(long long n, int world) {
    int g = sample_ctas(n);
    if (world > 1) {
        int cap = 64 / world;
        if (cap < 8) cap = 8;
        if (g > cap) g = cap;
    }
    return g;
}

constexpr double kZ = 5.5;  // bracket half-width in standard deviations of the sample-quantile estimate

inline double col_delta(long long n, int world) {
    double s_tot = (double)sample_ctas_w(n, world) * kSample * world;
    double delta = kZ * 0.5 / sqrt(s_tot);
    return delta < 2.0 / kSample ? 2.0 / kSample : delta;
}

// candidate capacities: the M band holds ~2*delta of the rows, L u R ~8*delta (each widened by the M band)
inline long long cap_for(long long n, double frac) {
    long long c = (long long)((double)n * frac * col_delta(n, 1) / 0.0054) + 32768;
    return c > n ? n : c;
}

}  // namespace

extern "C" {

// ---- fast column median / MAD -------------------------------------------------------------------------
size_t abcdls_mad_fast_workspace_bytes(int64_t n, int d) {
    const int G = sample_ctas(n);
    return carve_mad(nullptr, d, G, cap_for(n, 0.03), cap_for(n, 0.08)).total;
}

// Stage 1: sample every column -> local bracket estimates.  est_out/est_count describe the buffer a multi-GPU
// caller sums across ranks (then passes world to stage 2).
int abcdls_mad_fast_sample(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, void* ws,
                           size_t ws_bytes, double** est_out, size_t* est_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && ws && n >= kSample && d >= 1 && d <= kMaxD && ld >= n && world >= 1, "mad_fast_sample");
    const int G = sample_ctas(n);
    MadWs w = carve_mad(ws, d, G, cap_for(n, 0.03), cap_for(n, 0.08));
    if (ws_bytes < w.total) {
        h->err = "mad_fast workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const double delta = col_delta(n, world);
    const size_t smem = (size_t)2 * kSample * sizeof(double);
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_sample_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.cnt, 0, w.zero_bytes, st));
    const int Gw = sample_ctas_w(n, world);
    k_sample_cols<<<dim3(Gw, d), 512, smem, st>>>(ss, n, ld, 0x5EEDu, delta, w.est);
    ABCDLS_LAUNCH_CHECK(h);
    if (est_out) *est_out = w.est;
    if (est_count) *est_count = (size_t)Gw * d * 4;
    return ABCDLS_OK;
}

// Stage 2: brackets + the table pass.  counts_out/counts_count: buffer to sum across ranks afterwards
// (counts[d*4], cnt[d*2], histM[d*kNB], histU[d*kNB] are contiguous).
int abcdls_mad_fast_pass(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, void* ws,
                         size_t ws_bytes, int64_t** counts_out, size_t* counts_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && ws && n >= kSample && d >= 1 && d <= kMaxD && ld >= n && world >= 1, "mad_fast_pass");
    const int G = sample_ctas(n);
    const long long capM = cap_for(n, 0.03), capU = cap_for(n, 0.08);
    MadWs w = carve_mad(ws, d, G, capM, capU);
    if (ws_bytes < w.total) {
        h->err = "mad_fast workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_make_brackets<<<d, 32, 0, st>>>(w.est, sample_ctas_w(n, world), d, 1.0 / world, w.brk);
    ABCDLS_LAUNCH_CHECK(h);
    long long gx = ((n + kATile - 1) / kATile);
    long long mx = (long long)h->sm_count * 8 / d;
    if (mx < 1) mx = 1;
    if (gx > mx) gx = mx;
    k_bracket_pass<<<dim3((unsigned)gx, d), kAThreads, 0, st>>>(ss, n, ld, w.brk, w.counts, w.histM, w.candM, capM,
                                                                w.candU, capU, w.cnt);
    k_copy_u64<<<1, 128, 0, st>>>(w.gcnt, w.cnt, 2 * d);
    ABCDLS_LAUNCH_CHECK(h);
    if (counts_out) *counts_out = reinterpret_cast<int64_t*>(w.counts);
    if (counts_count) *counts_count = w.redA_count;   // counts | gcnt | histM, contiguous (zero padded)
    return ABCDLS_OK;
}

// Stage 3: finish median and MAD from the candidates, in four phases so that a multi-rank caller can exchange
// between them (single rank: abcdls_mad_fast_finish runs all four back to back):
//   phase 0: plan + locate bins + gather (median)      -> exchange: all-gather x0 = small block, x1 = counts
//   phase 1: sort the (gathered) values -> medians; plan MAD; histogram of |x - med| over the U candidates
//                                                       -> exchange: all-reduce x0 = histU
//   phase 2: locate bins + gather (MAD)                 -> exchange: all-gather x0 = small block, x1 = counts
//   phase 3: sort -> MAD, guards
// gath_small / gath_cnt: the all-gathered exchange block of the previous phase (NULL on a single rank).
// shared by the two-pass path and the one-pass path (fused.cuh)
static int refine_impl(abcdls_handle_t h, const MadWs& w, long long capM, long long capU, bool fused, int64_t n_global,
                       int d, int phase, int world, const double* gath_small, const int32_t* gath_cnt, double* med,
                       double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1, size_t* x1_count,
                       cudaStream_t st) {
    const long long k0 = (n_global - 1) / 2, k1 = n_global / 2;
    const size_t smem = (size_t)kSmallCap * sizeof(double);
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_refine_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto gx_for = [&](long long cap) {
        long long g = (cap + 256 * 8 - 1) / (256 * 8);
        long long mx = (long long)h->sm_count * 8 / d;
        if (mx < 1) mx = 1;
        return (unsigned)(g > mx ? mx : (g < 1 ? 1 : g));
    };
    const unsigned* gc = reinterpret_cast<const unsigned*>(gath_cnt);
    if (x0) *x0 = nullptr;
    if (x1) *x1 = nullptr;
    switch (phase) {
        case 0:
            launch_pdl(k_plan_median, dim3(1), dim3(64), 0, st, w.jobs, d, w.brk, w.counts, w.histM, w.candM, capM, w.cnt, w.gcnt, w.small,
                                            w.small_count, k0, k1, status_dev);
            launch_pdl(k_refine_pick, dim3(d), dim3(256), 0, st, w.jobs, status_dev);
            launch_pdl(k_refine_gather, dim3(dim3(gx_for(capM), d)), dim3(256), 0, st, w.jobs, capM);
            if (x0) { *x0 = w.small; *x0_count = (size_t)d * kSmallCap; }
            if (x1) { *x1 = w.small_count; *x1_count = (size_t)d; }
            break;
        case 1:
            launch_pdl(k_refine_final, dim3(d), dim3(1024), smem, st, w.jobs, gath_small, gc, gath_small ? world : 1, d, status_dev);
            launch_pdl(k_plan_mad, dim3(1), dim3(64), 0, st, w.jobs, d, w.brk, w.counts, w.histU, w.candU, capU, w.cnt, w.gcnt, w.small,
                                         w.small_count, k0, k1, med, w.guard, status_dev);
            launch_pdl(k_refine_hist, dim3(dim3(gx_for(capU), d)), dim3(256), 0, st, w.jobs + d, capU);
            if (x0) { *x0 = w.histU; *x0_count = w.redB_count; }
            break;
        case 2:
            launch_pdl(k_refine_pick, dim3(d), dim3(256), 0, st, w.jobs + d, status_dev);
            launch_pdl(k_refine_gather, dim3(dim3(gx_for(capU), d)), dim3(256), 0, st, w.jobs + d, capU);
            if (x0) { *x0 = w.small + (size_t)d * kSmallCap; *x0_count = (size_t)d * kSmallCap; }
            if (x1) { *x1 = w.small_count + d; *x1_count = (size_t)d; }
            break;
        default:
            launch_pdl(k_refine_final, dim3(d), dim3(1024), smem, st, w.jobs + d, gath_small, gc, gath_small ? world : 1, d, status_dev);
            launch_pdl(k_finish_mad, dim3(1), dim3(64), 0, st, w.jobs, d, w.guard, mad, status_dev);
            break;
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_mad_fast_refine(abcdls_handle_t h, int64_t n, int64_t n_global, int d, int phase, int world,
                           const double* gath_small, const int32_t* gath_cnt, void* ws, size_t ws_bytes, double* med,
                           double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1, size_t* x1_count,
                           void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && med && mad && status_dev && d >= 1 && d <= kMaxD && phase >= 0 && phase <= 3 &&
                            world >= 1 && world <= 64,
                     "mad_fast_refine");
    const int G = sample_ctas(n);
    const long long capM = cap_for(n, 0.03), capU = cap_for(n, 0.08);
    MadWs w = carve_mad(ws, d, G, capM, capU);
    if (ws_bytes < w.total) {
        h->err = "mad_fast workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    return refine_impl(h, w, capM, capU, false, n_global, d, phase, world, gath_small, gath_cnt, med, mad, status_dev, x0,
                       x0_count, x1, x1_count, (cudaStream_t)stream);
}

int abcdls_mad_fast_finish(abcdls_handle_t h, int64_t n, int64_t n_global, int d, void* ws, size_t ws_bytes,
                           double* med, double* mad, int* status_dev, void* stream) {
    for (int phase = 0; phase < 4; ++phase) {
        int rc = abcdls_mad_fast_refine(h, n, n_global, d, phase, 1, nullptr, nullptr, ws, ws_bytes, med, mad,
                                        status_dev, nullptr, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
    }
    return ABCDLS_OK;
}

// ---- fused distance + tolerance bracket ----------------------------------------------------------------
size_t abcdls_dist_fast_workspace_bytes(int64_t n, int64_t cap) {
    return carve_dist(nullptr, 64, n, cap).total;
}

int abcdls_dist_fast_sample(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* mad,
                            const double* target, int64_t nacc, int64_t n_global, int world, int64_t cap, void* ws,
                            size_t ws_bytes, double** est_out, size_t* est_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && mad && target && ws && n >= kSample && d >= 1 && d <= kMaxD && ld >= n,
                     "dist_fast_sample");
    const int G = 64;
    DistWs w = carve_dist(ws, G, n, cap);
    if (ws_bytes < w.total) {
        h->err = "dist_fast workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const double p = (double)nacc / (double)n_global;
    const double s_tot = (double)G * kSample * world;
    double delta = kZ * sqrt(p * (1.0 - p) / s_tot);
    const size_t smem = (size_t)kSample * sizeof(double);
    ABCDLS_CUDA(h, cudaMemsetAsync(w.cand_cnt, 0, w.zero_bytes, st));
    k_sample_dist<<<G, 512, smem, st>>>(ss, n, d, ld, mad, target, 0xD157u, p, delta, w.est);
    ABCDLS_LAUNCH_CHECK(h);
    if (est_out) *est_out = w.est;
    if (est_count) *est_count = (size_t)G * 2;
    return ABCDLS_OK;
}

// exact ds among the candidates in [lo, hi].  phase 0: plan + locate bins + gather (exchange: all-gather
// x0 = small block, x1 = count); phase 1: sort the (gathered) values -> ds_out.
int abcdls_dist_fast_refine(abcdls_handle_t h, int64_t n, int64_t nacc, int64_t cap, int phase, int world,
                            const double* gath_small, const int32_t* gath_cnt, const double* cand_dist, double* ds_out,
                            void* ws, size_t ws_bytes, int* status_dev, void** x0, size_t* x0_count, void** x1,
                            size_t* x1_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && cand_dist && ds_out && status_dev && (phase == 0 || phase == 1) && world >= 1 &&
                            world <= 64,
                     "dist_fast_refine");
    DistWs w = carve_dist(ws, 64, n, cap);
    if (ws_bytes < w.total) {
        h->err = "dist_fast workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)kSmallCap * sizeof(double);
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_refine_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (phase == 0) {
        k_plan_ds<<<1, 32, 0, st>>>(w.job, w.brk, w.n_lt, w.cand_cnt, w.gcand_cnt, w.hist, cand_dist, cap, w.small,
                                    w.small_count, nacc - 1, status_dev);
        k_refine_pick<<<1, 256, 0, st>>>(w.job, status_dev);
        long long gg = (cap + 2047) / 2048;
        if (gg > (long long)h->sm_count * 4) gg = (long long)h->sm_count * 4;
        k_refine_gather<<<dim3((unsigned)gg, 1), 256, 0, st>>>(w.job, cap);
        if (x0) { *x0 = w.small; *x0_count = (size_t)kSmallCap; }
        if (x1) { *x1 = w.small_count; *x1_count = 1; }
    } else {
        k_refine_final<<<1, 1024, smem, st>>>(w.job, gath_small, reinterpret_cast<const unsigned*>(gath_cnt),
                                              gath_small ? world : 1, 1, status_dev);
        k_copy_result<<<1, 32, 0, st>>>(w.job, ds_out);
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// dist may be NULL (the N-vector is then never materialised).  Outputs (row order): cand_row (global rows),
// cand_dist, n_cand (device u64).  ds_out (device double) = exact nacc-th smallest distance.
int abcdls_dist_fast_pass(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* mad,
                          const double* target, int64_t nacc, int world, int64_t row_offset, int64_t cap, double* dist,
                          int64_t* cand_row, double* cand_dist, int64_t* n_cand, double* ds_out, void* ws,
                          size_t ws_bytes, int* status_dev, int64_t** red_out, size_t* red_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && mad && target && cand_row && cand_dist && n_cand && ds_out && ws && status_dev,
                     "dist_fast_pass: null");
    ABCDLS_CHECK_ARG(h, n >= kSample && d >= 1 && d <= kMaxD && ld >= n && cap >= 1, "dist_fast_pass: shape");
    const int G = 64;
    DistWs w = carve_dist(ws, G, n, cap);
    if (ws_bytes < w.total) {
        h->err = "dist_fast workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_make_dist_bracket<<<1, 32, 0, st>>>(w.est, G, 1.0 / world, w.brk);
    const bool vec = ((reinterpret_cast<uintptr_t>(ss) & 15) == 0) && ((ld & 1) == 0) &&
                     (!dist || (reinterpret_cast<uintptr_t>(dist) & 15) == 0);
    const long long ntiles = n / kTRows;   // full 256-row tiles go through the TMA ring
    int nprod;
    long long rcap;
#define LAUNCH_LDG(V, W, GRID, ROW_LO, PROD0, RPW)                                                                 \
    k_dist_pass<V, W><<<GRID, kBThreads, 0, st>>>(ss, n, d, ld, mad, target, w.brk, dist, w.row_tmp, w.dist_tmp,    \
                                                  cap, w.cand_cnt, w.n_lt, w.hist, ROW_LO, PROD0, RPW, rcap,        \
                                                  w.rec_off, w.rec_cnt, w.wtot, w.nrec, status_dev)
    if (vec && ntiles >= kTGrid && !h->no_tma) {
        const long long tpc = tma_tiles_per_cta(ntiles);
        rcap = tma_recs_per_cta(ntiles);
        if (rcap < 2 * 1 + 2) rcap = 4;
        nprod = kTGrid + 8;
        const size_t smem = sizeof(TmaSmem);
        if (dist) {
            ABCDLS_ONCE(h, cudaFuncSetAttribute(k_dist_pass_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_dist_pass_tma<true><<<kTGrid, kTThreads, smem, st>>>(ss, ntiles, d, ld, mad, target, w.brk, dist, w.row_tmp,
                                                                  w.dist_tmp, cap, w.cand_cnt, w.n_lt, w.hist, tpc, rcap,
                                                                  w.rec_off, w.rec_cnt, w.wtot, w.nrec, status_dev);
        } else {
            ABCDLS_ONCE(h, cudaFuncSetAttribute(k_dist_pass_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_dist_pass_tma<false><<<kTGrid, kTThreads, smem, st>>>(ss, ntiles, d, ld, mad, target, w.brk, dist, w.row_tmp,
                                                                   w.dist_tmp, cap, w.cand_cnt, w.n_lt, w.hist, tpc, rcap,
                                                                   w.rec_off, w.rec_cnt, w.wtot, w.nrec, status_dev);
        }
        ABCDLS_LAUNCH_CHECK(h);
        // ragged tail (< 256 rows): one CTA of the load/store kernel, directory slots kTGrid..kTGrid+7
        if (dist) LAUNCH_LDG(true, true, 1, ntiles * kTRows, kTGrid, (long long)kBWarpTile);
        else LAUNCH_LDG(true, false, 1, ntiles * kTRows, kTGrid, (long long)kBWarpTile);
    } else {
        const long long rpw = rows_per_warp(n);
        rcap = recs_per_warp(n);
        nprod = kBWarps;
        if (vec && dist) LAUNCH_LDG(true, true, kBGrid, 0ll, 0, rpw);
        else if (vec) LAUNCH_LDG(true, false, kBGrid, 0ll, 0, rpw);
        else if (dist) LAUNCH_LDG(false, true, kBGrid, 0ll, 0, rpw);
        else LAUNCH_LDG(false, false, kBGrid, 0ll, 0, rpw);
    }
#undef LAUNCH_LDG
    ABCDLS_LAUNCH_CHECK(h);
    k_scan_tiles<<<1, 1024, 0, st>>>(w.wtot, nprod, w.wpos);
    k_order_cands<<<(nprod + 7) / 8, 256, 0, st>>>(w.rec_off, w.rec_cnt, w.nrec, w.wpos, nprod, rcap, w.row_tmp,
                                                   w.dist_tmp, cap, row_offset, (long long*)cand_row, cand_dist);
    ABCDLS_CUDA(h, cudaMemcpyAsync(n_cand, w.cand_cnt, 8, cudaMemcpyDeviceToDevice, st));
    ABCDLS_CUDA(h, cudaMemcpyAsync(w.gcand_cnt, w.cand_cnt, 8, cudaMemcpyDeviceToDevice, st));
    if (red_out) *red_out = reinterpret_cast<int64_t*>(w.gcand_cnt);
    if (red_count) *red_count = w.red_count;   // gcand_cnt | n_lt | hist: all-reduce(sum) across ranks
    if (world == 1) {
        for (int phase = 0; phase < 2; ++phase) {
            int rc = abcdls_dist_fast_refine(h, n, nacc, cap, phase, 1, nullptr, nullptr, cand_dist, ds_out, ws, ws_bytes,
                                             status_dev, nullptr, nullptr, nullptr, nullptr, stream);
            if (rc) return rc;
        }
    }
    return ABCDLS_OK;
}

}  // extern "C"

#include "fused.cuh"
#include "columns.cuh"
