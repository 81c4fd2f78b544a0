// fused.cuh — the ONE-pass front half of abc()/postpr(): column medians + MADs AND the tolerance cut from a single
// read of the N x d table (included at the end of bracket.cu; shares its brackets, refiners and workspace).
//
// R's data dependency (the MADs must be known before any scaled distance) is broken with interval
// reasoning instead of a second read:
//   1. the column samples that bracket median / MAD also give an estimate mu_j of every MAD with a relative
//      uncertainty eps (the width of the 5.5-sigma sample band);
//   2. a row sample evaluated with rho_j = 1/mu_j brackets the nacc-th smallest APPROXIMATE squared distance
//      q~(x) = sum_j ((x_j - t_j) rho_j)^2  from above: Q_hi;
//   3. the pass classifies every element against the column brackets (exactly as k_bracket_pass) and, from the same
//      shared-memory tile, marks every row with q~ <= thr = Q_hi ((1+eps)/(1-eps))^2 in a row bitmap;
//   4. once the exact MADs M_j are known the marked rows (~1.5 nacc) get R's exact distance, ds is selected among
//      them, and the result is PROVEN complete on the device:  |rho_j M_j - 1| <= eps for all j and
//      ds^2 (1+eps)^2 (1+eta) <= thr  imply that every unmarked row has dist > ds.  Otherwise FASTPATH_MISS is set
//      and the caller reruns with the two-pass path.  Exact or flagged, never approximate.
//
// Table streaming: TMA bulk copies (cp.async.bulk, 2 KB per column per 256-row tile) into an up-to-8-stage
// shared-memory ring (192 KB, one CTA per SM), 16 consumer warps.  Warp w classifies column w of the tile (8
// elements per lane from conflict-free 16-byte shared loads) and appends band candidates straight to 256-slot
// chunks in HBM (one global atomic per chunk, chunk fills in a directory).  For the distance each row is shared by
// two lanes (even / odd columns) and combined with one shuffle; the 16-bit ballot of the warp is one store into
// the row bitmap, so candidate rows come out in row order without any barrier, atomic or staging.

#include <cuda.h>

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// one 2-D tile (256 rows x d columns of the column-major table) per TMA instruction; rows beyond n arrive as NaN
__device__ __forceinline__ void tma_tile_g2s(void* dst_smem, const CUtensorMap* tmap, int row0, int col0,
                                             unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(row0), "r"(col0), "r"(smem_u32(bar))
        : "memory");
}

constexpr int kFRows = 256;
constexpr int kFNarrowD = 16;    // k_fused_pass: 256-row tiles, one column per consumer warp, 16-double candidate records
constexpr int kFMaxD = 32;       // k_fused_pass_wide (fused_wide.cuh): 128-row tiles, two columns per warp, 32-double records
inline int fused_tile_rows(int d) { return d > kFNarrowD ? 128 : kFRows; }
inline int fused_emit_pitch(int d) { return d > kFNarrowD ? 32 : kFNarrowD; }
constexpr int kFWarps = 16;
constexpr int kFConsumers = kFWarps * 32;
constexpr int kFThreads = kFConsumers + 32;
constexpr int kFMaxStages = 8;
constexpr int kFRingBytes = 160 * 1024;
constexpr int kFChunk = 256;
constexpr int kFEChunk = 64;     // candidate-row slots a warp allocates at a time
constexpr int kFG2 = 128;        // CTAs of the row sample, kSampleQ rows each (262144 rows in total)
constexpr int kSampleQ = 2048;
constexpr int kFHeader = 1024;   // bytes in front of the ring
constexpr int kFQCap = 16;       // per-lane queue slots (both bands share it; flushed when a lane holds more than 8)
constexpr int kFQWarpBytes = kFQCap * 256;                // per consumer warp
constexpr int kFQBytes = kFWarps * kFQWarpBytes;          // 64 KB

struct FusedPar {
    double rho[kFMaxD];   // 1 / (1.4826 * sampled MAD estimate)
    double tgt[kFMaxD];   // unscaled target
    double eps;           // relative uncertainty assumed for every rho_j * MAD_j
    double eta;           // rounding slack
    double qhi;           // sampled upper bracket of the nacc-th smallest q~
    double thr;           // candidate threshold on q~   (< 0: the one-pass path is not usable)
    int ok;
};

struct FusedSmem {
    unsigned long long full[kFMaxStages], empty[kFMaxStages];
    double2 sc[kFMaxD];   // (-target * rho, rho)
};
static_assert(sizeof(FusedSmem) <= kFHeader, "header");

inline int fused_stages(int d) {
    int s = kFRingBytes / (d * fused_tile_rows(d) * 8);
    return s > kFMaxStages ? kFMaxStages : s;
}

// scale estimates from the (rank-summed) column sample: est[g][j][{m_lo, m_hi, r_lo, r_hi}]
__global__ void k_make_fused_scale(const double* __restrict__ est, int G, int d, double inv_world,
                                   const double* __restrict__ target, FusedPar* par) {
    pdl_prologue();
    const int j = threadIdx.x;   // one warp
    double eps = 0.0;
    bool ok = true;
    if (j < d) {
        double a = 0.0, b = 0.0;
#pragma unroll 8
        for (int g = 0; g < G; ++g) {
            a += est[((size_t)g * d + j) * 4 + 2];
            b += est[((size_t)g * d + j) * 4 + 3];
        }
        a = a / G * inv_world;
        b = b / G * inv_world;
        const double mid = 0.5 * a + 0.5 * b;
        ok = (mid > 1e-150) && (mid < 1e150) && (a > 0.0);
        par->rho[j] = ok ? 1.0 / (1.4826 * mid) : 0.0;
        par->tgt[j] = target[j];
        eps = ok ? 1.25 * (b - a) / (b + a) + 1e-7 : 1.0;
        ok = ok && isfinite(target[j]);
    } else if (j < kFMaxD) {
        par->rho[j] = 0.0;
        par->tgt[j] = 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) eps = fmax(eps, __shfl_xor_sync(0xffffffffu, eps, o));
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (j == 0) {
        par->eps = eps;
        par->ok = (all_ok && eps < 0.2) ? 1 : 0;
        par->thr = -1.0;
        par->qhi = 0.0;
        par->eta = 1e-9;
    }
}

// q~ of a random sample of rows -> per-CTA estimate of the (p + delta) quantile.  grid G, 512 threads
__global__ void __launch_bounds__(512) k_sample_qdist(const double* __restrict__ ss, long long n, int d, long long ld,
                                                      const FusedPar* __restrict__ par, unsigned seed, double p,
                                                      double delta, double* __restrict__ est) {
    pdl_prologue();
    extern __shared__ double sm[];
    __shared__ double2 sc[kFMaxD];
    if (threadIdx.x < kFMaxD) sc[threadIdx.x] = make_double2(par->tgt[threadIdx.x], par->rho[threadIdx.x]);
    __syncthreads();
    for (int t = threadIdx.x; t < kSampleQ; t += blockDim.x) {
        const long long r = sample_row(seed, blockIdx.x, t, n);
        double q = 0.0;
        for (int j = 0; j < d; ++j) {
            const double a = (ss[(size_t)j * ld + r] - sc[j].x) * sc[j].y;
            q = fma(a, a, q);
        }
        sm[t] = (q == q) ? q : __longlong_as_double(0x7ff0000000000000ll);   // NaN rows sort last
    }
    __syncthreads();
    bitonic_sort(sm, kSampleQ);
    if (threadIdx.x == 0) {
        int ih = (int)ceil((p + delta) * kSampleQ) + 1;
        ih = max(0, min(kSampleQ - 1, ih));
        est[blockIdx.x] = sm[ih];
    }
}

__global__ void k_make_fused_thr(const double* __restrict__ est, int G, double inv_world, int d, FusedPar* par) {
    pdl_prologue();
    if (threadIdx.x || blockIdx.x) return;
    double hi = 0.0;
    for (int g = 0; g < G; ++g) hi += est[g];
    hi = hi / G * inv_world;
    double amax = 0.0;
    for (int j = 0; j < d; ++j) amax = fmax(amax, fabs(par->tgt[j] * par->rho[j]));
    const double eps = par->eps;
    // rounding of R's (x/m - t/m)^2 terms: absolute error ~ 2^-52 |t/m| per term against distances ~ sqrt(hi)
    const double eta = 1e-9 + 1e-14 * (1.0 + amax / sqrt(fmax(hi, 1e-300)));
    const bool ok = par->ok && hi > 0.0 && isfinite(hi) && eta < 1e-3;
    const double f = (1.0 + eps) / (1.0 - eps);
    par->qhi = hi;
    par->eta = eta;
    par->thr = ok ? hi * f * f * (1.0 + eta) * (1.0 + eta) : -1.0;
    if (!ok) par->ok = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Second-stage column sample: COUNTING instead of sorting.  The sorted first-stage sample (k_sample_cols, now only
// 8 CTAs per column) locates every bracket to +-5.5 sigma_1; a much larger sample (one 128-byte line of 16
// consecutive rows out of every 2^shift lines: stratified, one jittered line per stratum, so it streams like a
// strided read) is then only CLASSIFIED against those windows - three 1024-bin histograms per column (the median
// window, and |x - c| in the MAD window left / right of the centre c) plus the counts outside them.  From the
// cumulative counts k_fine_brackets reads new brackets at +-5.5 sigma_2 (bin edges, conservative by one bin), which
// are sqrt(S2 / S1) times narrower: the pass keeps / writes / the refinement re-reads that much less, and the
// relative uncertainty eps of the MAD estimates (hence the number of emitted candidate rows) shrinks with it.
// A wrong bracket can only cost a FASTPATH_MISS: the pass still counts exactly and proves completeness.
constexpr int kFineNB = 1024;
constexpr int kFineU32 = 3 * kFineNB + 4;   // per column: M window | left wing | right wing | {below, farL, farR, taken}
static_assert(kFineU32 % 2 == 0, "two 32-bit counts travel as one int64 of the all-reduce");

__global__ void __launch_bounds__(512) k_sample_fine(const double* __restrict__ ss, long long n, long long ld,
                                                     const ColBrk* __restrict__ brk, long long L, int shift, int rs,
                                                     unsigned seed, unsigned* __restrict__ out) {
    pdl_prologue();
    __shared__ unsigned h[3 * kFineNB];
    __shared__ unsigned cn[4];
    const int j = blockIdx.y;
    for (int t = threadIdx.x; t < 3 * kFineNB; t += blockDim.x) h[t] = 0u;
    if (threadIdx.x < 4) cn[threadIdx.x] = 0u;
    __syncthreads();
    const ColBrk B = brk[j];
    const double c = B.c;
    // window coordinates as ONE fma each: bin = floor(coordinate); < 0 below the window, >= kFineNB above it.  The
    // edges k_fine_brackets reasons about are these integer coordinates, so both kernels agree by construction.
    const double sM = B.r0 > 0.0 ? (double)kFineNB / (2.0 * B.r0) : 0.0, oM = B.r0 * sM;
    const double sW = B.r2 > B.r1 ? (double)kFineNB / (B.r2 - B.r1) : 0.0, oW = -B.r1 * sW;
    const double* __restrict__ col = ss + (size_t)j * ld + 2 * (threadIdx.x & 7);   // 8 lanes x 16 bytes = one line
    const int hw = threadIdx.x >> 3;
    const long long nhw = (long long)gridDim.x * (blockDim.x >> 3);   // quarter-warps working on this column
    // a stratum is 2^(shift + rs) lines, of which a run of 2^rs consecutive ones (run-aligned, jittered) is read: the
    // same sample size as one line in 2^shift, in 2^rs times fewer DRAM pages
    const unsigned run = 1u << rs, jmask = ((1u << (shift + rs)) - 1u) & ~(run - 1u);
    const unsigned salt = hash3(seed, (unsigned)j, 0x51u);
    const unsigned hM = smem_u32(h), hL = hM + kFineNB * 4;
    unsigned below = 0, far = 0, farneg = 0, taken = 0;
    for (long long q = (long long)blockIdx.x * (blockDim.x >> 3) + hw; 4 * q < L; q += nhw) {   // L is a multiple of 4
        double x[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long s = 4 * q + u, stratum = s >> rs;
            unsigned hsh = ((unsigned)stratum ^ salt) * 0x9E3779B1u;      // jitter only: a multiplicative hash will do
            hsh ^= hsh >> 15;
            hsh *= 0x85EBCA77u;
            const long long line = (stratum << (shift + rs)) + (long long)(((hsh >> 12) & jmask) + ((unsigned)s & (run - 1u)));
            const double2 v = ld_stream2(col + line * 16);
            x[2 * u] = v.x;
            x[2 * u + 1] = v.y;
        }
        taken += 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double sg = __dsub_rn(x[u], c);
            const int bm = __double2int_rd(fma(sg, sM, oM));
            const int bw = __double2int_rd(fma(fabs(sg), sW, oW));
            const unsigned neg = (unsigned)__double2hiint(sg) >> 31;
            below += (unsigned)bm >> 31;
            const unsigned isfar = bw >= kFineNB ? 1u : 0u;
            far += isfar;
            farneg += isfar & neg;
            const unsigned inM = (unsigned)bm < (unsigned)kFineNB ? 1u : 0u, inW = (unsigned)bw < (unsigned)kFineNB ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred pm, pw;\n\t"
                "setp.ne.u32 pm, %0, 0;\n\t"
                "setp.ne.u32 pw, %1, 0;\n\t"
                "@pm red.shared.add.u32 [%2], 1;\n\t"
                "@pw red.shared.add.u32 [%3], 1;\n\t"
                "}" ::"r"(inM),
                "r"(inW), "r"(hM + (unsigned)bm * 4u), "r"(hL + ((neg ^ 1u) * kFineNB + (unsigned)bw) * 4u)
                : "memory");
        }
    }
    below = warp_sum(below);
    far = warp_sum(far);
    farneg = warp_sum(farneg);
    taken = warp_sum(taken);
    if ((threadIdx.x & 31) == 0) {
        if (below) atomicAdd(&cn[0], below);
        if (farneg) atomicAdd(&cn[1], farneg);
        if (far - farneg) atomicAdd(&cn[2], far - farneg);
        if (taken) atomicAdd(&cn[3], taken);
    }
    __syncthreads();
    unsigned* __restrict__ o = out + (size_t)j * kFineU32;
    for (int t = threadIdx.x; t < 3 * kFineNB; t += blockDim.x)
        if (h[t]) atomicAdd(&o[t], h[t]);
    if (threadIdx.x < 4 && cn[threadIdx.x]) atomicAdd(&o[3 * kFineNB + threadIdx.x], cn[threadIdx.x]);
}

// One CTA per column: cumulative counts of the three histograms (summed over ranks) -> refined estimates
// est[j] = {m_lo, m_hi, r_lo, r_hi} in the layout k_make_brackets / k_make_fused_scale average (with G = 1); est comes
// in holding the first-stage means and keeps them for a column whose refinement is not usable (degenerate window,
// a crossing outside the window).  r_lo / r_hi bracket the median of |x - c2| around the NEW centre c2 = (m_lo + m_hi)/2:
// #(|x - c2| <= r) = S - #(x - c > r + D) - #(c - x > r - D), D = c2 - c, each term bounded from the wing histograms.
__global__ void __launch_bounds__(256) k_fine_brackets(const unsigned* __restrict__ fine,
                                                       const ColBrk* __restrict__ brk, double z,
                                                       double* __restrict__ est) {
    pdl_prologue();
    __shared__ long long cum[kFineNB + 1];     // cum[k]  = # M-window samples in bins < k
    __shared__ long long suf[2][kFineNB + 1];  // suf[k]  = # samples of that side with radius >= edge k (incl. far)
    __shared__ long long part[3][8];
    __shared__ int kres[4];
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned* __restrict__ f = fine + (size_t)j * kFineU32;
    const long long below = f[3 * kFineNB + 0], farL = f[3 * kFineNB + 1], farR = f[3 * kFineNB + 2];
    const long long S = f[3 * kFineNB + 3];
    // block scans: thread t owns bins 4t .. 4t+3 of each histogram
    long long loc[3][4], tot[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        tot[a] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            loc[a][q] = f[a * kFineNB + 4 * tid + q];
            tot[a] += loc[a][q];
        }
    }
    long long incl[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        incl[a] = tot[a];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl[a], o);
            if (lane >= o) incl[a] += t;
        }
        if (lane == 31) part[a][wid] = incl[a];
    }
    if (tid < 4) kres[tid] = -1;
    __syncthreads();
    long long allw[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        long long before = 0, all = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < wid) before += part[a][w];
            all += part[a][w];
        }
        incl[a] += before;
        allw[a] = all;
    }
    {
        long long e = incl[0] - tot[0];                    // exclusive prefix of the M window
#pragma unroll
        for (int q = 0; q < 4; ++q) { cum[4 * tid + q] = e; e += loc[0][q]; }
        if (tid == 255) cum[kFineNB] = e;
#pragma unroll
        for (int a = 1; a < 3; ++a) {
            long long g = allw[a] - (incl[a] - tot[a]) + (a == 1 ? farL : farR);   // radius >= edge 4 tid
#pragma unroll
            for (int q = 0; q < 4; ++q) { suf[a - 1][4 * tid + q] = g; g -= loc[a][q]; }
            if (tid == 255) suf[a - 1][kFineNB] = g;       // = far
        }
    }
    __syncthreads();
    const ColBrk B = brk[j];
    const bool usable = S >= 65536 && B.r0 > 0.0 && B.r2 > B.r1;
    const double dS = (double)S;
    const double delta = z * 0.5 / sqrt(fmax(dS, 1.0));
    const double tlo = (0.5 - delta) * dS, thi = (0.5 + delta) * dS;
    // median: A(k) = # samples below edge k of the M window
    int klo = -1, khi = 1 << 30;
    for (int k = tid; k <= kFineNB; k += 256) {
        const double A = (double)(below + cum[k]);
        if (A <= tlo) klo = max(klo, k);
        if (A >= thi) khi = min(khi, k);
    }
    klo = __reduce_max_sync(0xffffffffu, klo);
    khi = __reduce_min_sync(0xffffffffu, khi);
    if (lane == 0) {
        atomicMax(&kres[0], klo);
        atomicMax(&kres[1], (1 << 30) - khi);
    }
    __syncthreads();
    klo = kres[0];
    khi = (1 << 30) - kres[1];
    const bool med_ok = usable && klo >= 0 && khi <= kFineNB && khi > klo;
    const double wM = 2.0 * B.r0 / kFineNB, wW = (B.r2 - B.r1) / kFineNB;
    const double m_lo = (B.c - B.r0) + klo * wM, m_hi = (B.c - B.r0) + khi * wM;
    const double c2 = 0.5 * m_lo + 0.5 * m_hi;
    const double t = med_ok ? (c2 - B.c) / wW : 0.0;       // centre shift in wing bins
    int rlo = -1, rhi = 1 << 30;
    for (int k = tid; k <= kFineNB; k += 256) {
        // right side: x - c > r_k + D  <=>  wing coordinate > k + t;  left side: c - x > r_k - D  <=>  > k - t
        long long lb = 0, ub = 0;
        bool ub_ok = true;
#pragma unroll
        for (int sd = 0; sd < 2; ++sd) {
            const double pos = sd ? (double)k + t : (double)k - t;      // sd 0: left wing, 1: right wing
            int el = (int)ceil(pos) + 1, eu = (int)floor(pos) - 1;
            el = el < 0 ? 0 : (el > kFineNB ? kFineNB : el);
            lb += suf[sd][el];
            if (eu < 0) ub_ok = false;
            eu = eu < 0 ? 0 : (eu > kFineNB ? kFineNB : eu);
            ub += suf[sd][eu];
        }
        const double in_up = (double)(S - lb);                         // upper bound of #(|x - c2| <= r_k)
        const double in_dn = ub_ok ? (double)(S - ub) : -1.0;          // lower bound
        if (in_up <= tlo) rlo = max(rlo, k);
        if (in_dn >= thi) rhi = min(rhi, k);
    }
    rlo = __reduce_max_sync(0xffffffffu, rlo);
    rhi = __reduce_min_sync(0xffffffffu, rhi);
    if (lane == 0) {
        atomicMax(&kres[2], rlo);
        atomicMax(&kres[3], (1 << 30) - rhi);
    }
    __syncthreads();
    rlo = kres[2];
    rhi = (1 << 30) - kres[3];
    if (tid == 0 && med_ok && rlo >= 0 && rhi <= kFineNB && rhi > rlo) {
        est[j * 4 + 0] = m_lo;
        est[j * 4 + 1] = m_hi;
        est[j * 4 + 2] = B.r1 + rlo * wW;
        est[j * 4 + 3] = B.r1 + rhi * wW;
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFThreads, 1)
k_fused_pass(const __grid_constant__ CUtensorMap tmap, long long n, long long ntiles, int d, int nstages,
             const ColBrk* __restrict__ brk, const FusedPar* __restrict__ par, unsigned long long* __restrict__ counts,
             double* __restrict__ candM, long long capM, double* __restrict__ candU, long long capU,
             unsigned long long* __restrict__ cnt, unsigned short* __restrict__ mask16, double* __restrict__ cand_e,
             long long ecap, int* __restrict__ erow, unsigned long long* __restrict__ ealloc, long long tpc,
             int chunkM, int chunkU, int* status) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
    double* ring = reinterpret_cast<double*>(smem_raw + kFHeader);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kFMaxD)   // a = x * rho - t * rho as one fma
        sm.sc[threadIdx.x] = make_double2(-par->tgt[threadIdx.x] * par->rho[threadIdx.x], par->rho[threadIdx.x]);
    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < nstages; ++s2) {
            mbar_init(&sm.full[s2], 1);
            mbar_init(&sm.empty[s2], kFWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const long long t_begin = (long long)blockIdx.x * tpc, t_end = min(ntiles, t_begin + tpc);
    const size_t stage_doubles = (size_t)d * kFRows;
    if (wid == kFWarps) {
        // ===== producer: one elected lane, ONE 2-D TMA tile copy per stage =====
        if (lane == 0) {
            int st = 0;
            unsigned ph = 0;
            for (long long tile = t_begin; tile < t_end; ++tile) {
                mbar_wait(&sm.empty[st], ph ^ 1u);
                mbar_expect_tx(&sm.full[st], (unsigned)(stage_doubles * 8));   // out-of-bounds rows count too
                tma_tile_g2s(ring + (size_t)st * stage_doubles, &tmap, (int)(tile * kFRows), 0, &sm.full[st]);
                if (++st == nstages) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ===== consumers =====
    const int j = wid;                  // column classified by this warp
    const bool has_col = j < d;
    const int jc = has_col ? j : 0;
    const ColBrk B = brk[jc];
    const double c = B.c, r0 = B.r0, mr1 = -B.r1;
    // U band as an integer test on the high word of u = fl(|x - c| - r1): u >= 0 and hi(u) <= hi(fl(r2 - r1)).
    // fl(a - b) is monotone and never rounds a non-zero difference to zero, so the sign of u is exact (y < r1 <=> u < 0)
    // and every y in [r1, r2] passes; the few values just above r2 that share the top word only widen the band.
    const unsigned hiW = (unsigned)__double2hiint(__dsub_rn(B.r2, B.r1));
    const double thr = par->thr;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    // Band candidates (either band) go to a per-LANE queue in shared memory ([slot][lane]: conflict free, no votes, no
    // branches: one predicated store + one predicated pointer bump per element).  When some lane could overflow in
    // the next tile the warp flushes: entries are re-classified (M: |x - c| <= r0, else U), two scans, two global
    // atomics, dense copies.
    unsigned char* qbase = smem_raw + kFHeader + (size_t)nstages * stage_doubles * 8 + (size_t)wid * kFQWarpBytes;
    const unsigned q0 = smem_u32(qbase) + lane * 8;                        // my queue: slot s at q0 + s * 256
    unsigned pq = q0;
    const unsigned qlim = q0 + (kFQCap - 8) * 256;
    double* __restrict__ gM = candM + (size_t)jc * capM;
    double* __restrict__ gU = candU + (size_t)jc * capU;
    unsigned long long* cntM = &cnt[jc * 2 + 0];
    unsigned long long* cntU = &cnt[jc * 2 + 1];
    long long posM = 0, endM = 0, posU = 0, endU = 0;   // this warp's reserved chunk of each global list (warp uniform)
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned c_neg = 0, c_in = 0;                // per lane: x < c, |x - c| < r1
    unsigned eoff = 0;                           // candidate-row emission: next free slot of this warp's 64-slot chunk
    int eroom = 0;
    bool bad = false;
    const int row = (wid << 4) | (lane & 15);    // row of the tile this lane helps with
    const int half = lane >> 4;                  // 0: even columns, 1: odd columns
    auto flush = [&]() {
        const int fill = (int)((pq - q0) >> 8);
        const int mx = __reduce_max_sync(0xffffffffu, fill);
        // my entries, split by band: packed (nM | nU << 16) inclusive scan over the lanes
        int mine = 0;
        for (int t = 0; t < mx; ++t) {
            double v;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(q0 + t * 256));
            if (t < fill) mine += (fabs(__dsub_rn(v, c)) <= r0) ? 1 : 0x10000;
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        // Space in the global lists comes from chunks this warp reserved earlier (column j of this CTA belongs to this
        // warp alone): a flush pays the ~1 us round trip of a global atomic only when a chunk runs out.  A flush stalls
        // its warp, the warp holds its ring stage, and five stages are ~2.5 us of work - with two atomics per flush the
        // ring kept draining.  What is left of a chunk is padded with NaN, which every reader of the lists skips.
        const long long nM = tot & 0xffff, nU = tot >> 16;
        if (nM > endM - posM) {
            for (long long i = posM + lane; i < endM; i += 32)
                if (i < capM) gM[i] = qnan;
            const long long take = nM > chunkM ? nM : chunkM;
            unsigned long long b = 0ull;
            if (lane == 0) b = atomicAdd(cntM, (unsigned long long)take);
            posM = (long long)__shfl_sync(0xffffffffu, b, 0);
            endM = posM + take;
        }
        if (nU > endU - posU) {
            for (long long i = posU + lane; i < endU; i += 32)
                if (i < capU) gU[i] = qnan;
            const long long take = nU > chunkU ? nU : chunkU;
            unsigned long long b = 0ull;
            if (lane == 0) b = atomicAdd(cntU, (unsigned long long)take);
            posU = (long long)__shfl_sync(0xffffffffu, b, 0);
            endU = posU + take;
        }
        // branch-free copy-out: one select of the destination per entry, no divergence between the two bands; the
        // capacity test runs per entry only when this warp's chunk sticks out of a list (which is then a flagged miss)
        double* pM = gM + posM + ((incl - mine) & 0xffff);
        double* pU = gU + posU + ((incl - mine) >> 16);
        posM += nM;
        posU += nU;
        const bool roomy = endM <= capM && endU <= capU;
        for (int t = 0; t < mx; ++t) {
            double v;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(q0 + t * 256));
            if (t < fill) {
                const bool isM = fabs(__dsub_rn(v, c)) <= r0;
                double* dst = isM ? pM : pU;
                if (roomy || (isM ? (pM - gM) < capM : (pU - gU) < capU)) *dst = v;
                pM += isM ? 1 : 0;
                pU += isM ? 0 : 1;
            }
        }
        pq = q0;
        __syncwarp();
    };

    int st = 0;
    unsigned ph = 0;
    unsigned short* mptr = mask16 + t_begin * kFWarps + wid;
    const long long tail_tile = (n % kFRows) ? ntiles - 1 : -1;   // the ragged tile (rows >= tail_rows hold NaN)
    const int tail_rows = (int)(n % kFRows);
    for (long long tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&sm.full[st], ph);
        const double* __restrict__ stage = ring + (size_t)st * stage_doubles;
        double q = 0.0;
        // One element of the column strip: two subtractions, two sign-bit counters, one integer and one fp64 compare,
        // a predicated store + predicated pointer bump into this lane's queue.  (No "memory" clobber: the queue is only
        // touched by volatile asm, which keeps its own order, so the compiler may interleave the distance math.)
#define ABCDLS_CLASSIFY(xv)                                                                          \
        {                                                                                            \
            const double sgn = __dsub_rn((xv), c);                                                   \
            const double uu = __dadd_rn(fabs(sgn), mr1); /* |x - c| - r1 */                          \
            const unsigned hu = (unsigned)__double2hiint(uu);                                        \
            c_neg += (unsigned)__double2hiint(sgn) >> 31;                                            \
            c_in += hu >> 31;                                                                        \
            const bool cand = (fabs(sgn) <= r0) || (hu <= hiW); /* M band or U band */               \
            asm volatile(                                                                            \
                "{\n\t"                                                                              \
                ".reg .pred pc;\n\t"                                                                 \
                "setp.ne.u32 pc, %1, 0;\n\t"                                                         \
                "@pc st.shared.f64 [%0], %2;\n\t"                                                    \
                "@pc add.u32 %0, %0, 256;\n\t"                                                       \
                "}"                                                                                  \
                : "+r"(pq)                                                                           \
                : "r"((unsigned)cand), "d"(xv));                                                     \
        }
        if (d == kFNarrowD) {
            // every warp owns a column: ONE straight-line block, so that the classification chains and the distance
            // FMAs of the tile overlap in the pipeline
            if (__any_sync(0xffffffffu, pq > qlim)) flush();
            const double* __restrict__ colp = stage + j * kFRows + 2 * lane;
            const double* __restrict__ xp = stage + half * kFRows + row;
            const double2* __restrict__ sp = sm.sc + half;      // (-t rho, rho) of this lane's columns: broadcast loads
            double x[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 v = *reinterpret_cast<const double2*>(colp + 64 * u);
                x[2 * u] = v.x;
                x[2 * u + 1] = v.y;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double2 s2 = sp[2 * u];
                const double a = fma(xp[2 * u * kFRows], s2.y, s2.x);
                q = fma(a, a, q);
                ABCDLS_CLASSIFY(x[u]);
            }
        } else {
            if (has_col) {
                if (__any_sync(0xffffffffu, pq > qlim)) flush();
                const double* __restrict__ colp = stage + j * kFRows + 2 * lane;
                double x[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double2 v = *reinterpret_cast<const double2*>(colp + 64 * u);
                    x[2 * u] = v.x;
                    x[2 * u + 1] = v.y;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) ABCDLS_CLASSIFY(x[u]);
            }
            // approximate squared distance of row `row`: this lane sums the columns of its parity
            for (int cc = half; cc < d; cc += 2) {
                const double2 s2 = sm.sc[cc];
                const double a = fma(stage[cc * kFRows + row], s2.y, s2.x);
                q = fma(a, a, q);
            }
        }
#undef ABCDLS_CLASSIFY
        q += __shfl_xor_sync(0xffffffffu, q, 16);
        bad |= !(q < inf) && (tile != tail_tile || row < tail_rows);
        const unsigned bal = __ballot_sync(0xffffffffu, q <= thr);
        const unsigned m16 = bal & 0xffffu;
        if (lane == 0) *mptr = (unsigned short)m16;
        mptr += kFWarps;
        if (m16) {
            // candidate rows leave the SM right here (statistics + row number, to a slot of this warp's chunk), so
            // that their exact distance later needs no second, scattered read of the table
            const int k = __popc(m16);
            eroom -= k;
            if (eroom < 0) {
                unsigned nb = 0;
                if (lane == 0) nb = (unsigned)atomicAdd(ealloc, (unsigned long long)kFEChunk);
                eoff = __shfl_sync(0xffffffffu, nb, 0);
                eroom = kFEChunk - k;
            }
            const int r16 = lane & 15;
            if ((m16 >> r16) & 1u) {
                const unsigned slot = eoff + __popc(m16 & ((1u << r16) - 1u));
                if ((long long)slot < ecap) {
                    // row-major block: 128 bytes per candidate; this lane writes columns [8 half, 8 half + 8)
                    if (half == 0) erow[slot] = (int)(tile * kFRows + row);
                    double* dst = cand_e + (size_t)slot * kFNarrowD + 8 * half;
                    const double* src = stage + (size_t)(8 * half) * kFRows + row;
#pragma unroll
                    for (int q2 = 0; q2 < 4; ++q2) {
                        const int c0 = 8 * half + 2 * q2;
                        if (c0 < d) {
                            double2 v;
                            v.x = src[(2 * q2) * kFRows];
                            v.y = (c0 + 1 < d) ? src[(2 * q2 + 1) * kFRows] : 0.0;
                            *reinterpret_cast<double2*>(dst + 2 * q2) = v;
                        }
                    }
                }
            }
            eoff += k;
            __syncwarp();
        }
        if (lane == 0) mbar_arrive(&sm.empty[st]);   // every lane is done with the stage
        if (++st == nstages) { st = 0; ph ^= 1u; }
    }
    if (has_col) {
        flush();
        for (long long i = posM + lane; i < endM; i += 32)
            if (i < capM) gM[i] = qnan;
        for (long long i = posU + lane; i < endU; i += 32)
            if (i < capU) gU[i] = qnan;
        c_neg = warp_sum(c_neg);
        c_in = warp_sum(c_in);
        if (lane == 0) {
            // counts[4j]: x < c for now; k_fused_histM subtracts the M-band values below c => "strictly below the band"
            if (c_neg) atomicAdd(&counts[j * 4 + 0], (unsigned long long)c_neg);
            if (c_in) atomicAdd(&counts[j * 4 + 2], (unsigned long long)c_in);
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, ABCDLS_S_NONFINITE);
}

#include "fused_wide.cuh"

// Fine histogram of the M band of every column from its (dense) candidates, after the pass: the pass itself stays
// free of the binning arithmetic and of 16M global reductions.  Also turns counts[4j] from "x < c" into "strictly
// below the M band" by subtracting the band values below c (same predicate as the pass: sign of fl(x - c)).
// grid (gx, d)
__global__ void __launch_bounds__(256) k_fused_histM(const double* __restrict__ candM, long long capM,
                                                     const unsigned long long* __restrict__ cnt,
                                                     const ColBrk* __restrict__ brk,
                                                     unsigned long long* __restrict__ histM,
                                                     unsigned long long* __restrict__ counts) {
    pdl_prologue();
    __shared__ unsigned hist[kNB];
    __shared__ unsigned negs;
    const int j = blockIdx.y;
    for (int t = threadIdx.x; t < kNB; t += blockDim.x) hist[t] = 0u;
    if (threadIdx.x == 0) negs = 0u;
    __syncthreads();
    long long nv = (long long)cnt[j * 2 + 0];
    if (nv > capM) nv = capM;
    const double lo = brk[j].loM, scale = brk[j].scaleM, c = brk[j].c;
    const double* __restrict__ src = candM + (size_t)j * capM;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned neg = 0;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nv; i0 += 4 * stride) {
        double xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * stride;
            xv[u] = i < nv ? ld_stream(src + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * stride >= nv || xv[u] != xv[u]) continue;      // NaN: padding of a reserved chunk
            atomicAdd(&hist[bin_of(xv[u], lo, scale)], 1u);
            neg += (unsigned)__double2hiint(__dsub_rn(xv[u], c)) >> 31;
        }
    }
    neg = warp_sum(neg);
    if ((threadIdx.x & 31) == 0 && neg) atomicAdd(&negs, neg);
    __syncthreads();
    for (int t = threadIdx.x; t < kNB; t += blockDim.x)
        if (hist[t]) atomicAdd(&histM[(size_t)j * kNB + t], (unsigned long long)hist[t]);
    if (threadIdx.x == 0 && negs) atomicAdd(&counts[j * 4 + 0], 0ull - (unsigned long long)negs);
}

// ---------------------------------------------------------------------------------------------------------------
// row bitmap -> ascending candidate rows.  Word w covers rows 32w .. 32w+31.
constexpr int kMWords = 8;   // words per thread
__global__ void __launch_bounds__(256) k_mask_count(const unsigned* __restrict__ bitmap, long long nwords,
                                                    int* __restrict__ tile_cnt, unsigned long long* total) {
    pdl_prologue();
    __shared__ int sh[33];
    const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * kMWords;
    int c = 0;
#pragma unroll
    for (int t = 0; t < kMWords; ++t)
        if (base + t < nwords) c += __popc(bitmap[base + t]);
    int tot;
    block_exclusive_scan(c, sh, &tot);
    if (threadIdx.x == 0) {
        tile_cnt[blockIdx.x] = tot;
        if (tot) atomicAdd(total, (unsigned long long)tot);
    }
}

__global__ void __launch_bounds__(256) k_mask_write(const unsigned* __restrict__ bitmap, long long nwords,
                                                    const long long* __restrict__ tile_off, long long row_offset,
                                                    long long cap, long long* __restrict__ rows_out,
                                                    int* __restrict__ wprefix,
                                                    const unsigned long long* __restrict__ total,
                                                    long long* __restrict__ n_out, int* status) {
    pdl_prologue();
    __shared__ int sh[33];
    const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * kMWords;
    unsigned wv[kMWords];
    int c = 0;
#pragma unroll
    for (int t = 0; t < kMWords; ++t) {
        wv[t] = (base + t < nwords) ? bitmap[base + t] : 0u;
        c += __popc(wv[t]);
    }
    int tot;
    const int ex = block_exclusive_scan(c, sh, &tot);
    long long pos = tile_off[blockIdx.x] + ex;
#pragma unroll
    for (int t = 0; t < kMWords; ++t) {
        unsigned m = wv[t];
        if (base + t < nwords) wprefix[base + t] = (int)pos;   // candidates before the first row of this word
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (pos < cap) rows_out[pos] = row_offset + (base + t) * 32 + b;
            ++pos;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const long long tt = (long long)*total;
        *n_out = tt < cap ? tt : cap;
        if (tt > cap) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    }
}

// R's exact distance of the emitted candidate rows (same arithmetic as k_dist_pass).  One thread per slot: coalesced
// reads of the compact row-major block (16 doubles per slot); the position of the row in the ascending candidate list comes from the bitmap
// (prefix of its word + bits below it), where the distance and the slot number are scattered to.
__global__ void __launch_bounds__(256) k_cand_exact(const double* __restrict__ cand_e, long long ecap,
                                                    const int* __restrict__ erow,
                                                    const unsigned long long* __restrict__ ealloc, int d, int pitch,
                                                    const double* __restrict__ mad, const double* __restrict__ target,
                                                    const unsigned* __restrict__ bitmap, const int* __restrict__ wprefix,
                                                    long long cap, double* __restrict__ cand_dist,
                                                    long long* __restrict__ perm, int* status) {
    pdl_prologue();
    __shared__ ColScale S;
    load_scale(S, mad, target, d);
    __syncthreads();
    long long ns = (long long)*ealloc;
    if (blockIdx.x == 0 && threadIdx.x == 0 && ns > ecap) atomicOr(status, ABCDLS_S_FASTPATH_MISS);   // rows were lost
    if (ns > ecap) ns = ecap;
    bool nonfinite = false;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long sl = (long long)blockIdx.x * blockDim.x + threadIdx.x; sl < ns; sl += stride) {
        const int r = erow[sl];
        if (r < 0) continue;   // unused tail of a chunk
        double s0 = 0.0;
        const double* __restrict__ rowp = cand_e + (size_t)sl * pitch;
        for (int j = 0; j < d; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(rowp + j);
            const double a = __dsub_rn(scaled_value(v.x, S.dr[j]), S.tgt[j]);
            s0 = __dadd_rn(s0, __dmul_rn(a, a));
            if (j + 1 < d) {
                const double b = __dsub_rn(scaled_value(v.y, S.dr[j + 1]), S.tgt[j + 1]);
                s0 = __dadd_rn(s0, __dmul_rn(b, b));
            }
        }
        const double dd = __dsqrt_rn(s0);
        nonfinite |= !isfinite(dd);
        const int word = r >> 5;
        const long long pos = (long long)wprefix[word] + __popc(bitmap[word] & ((1u << (r & 31)) - 1u));
        if (pos < cap) {
            cand_dist[pos] = dd;
            perm[pos] = sl;
        }
    }
    if (__any_sync(0xffffffffu, nonfinite) && (threadIdx.x & 31) == 0) atomicOr(status, ABCDLS_S_NONFINITE);
}

// ds = the (nacc - 1)-th (0-based, over all ranks) smallest exact candidate distance, with the histogram -> bin ->
// gather -> sort machinery of the brackets instead of six radix passes.  The candidate distances crowd towards the
// top (count(dist <= r) ~ r^d), so bins are uniform in (dist / hi)^(2^k), 2^k >= d: a monotone map that spreads them
// evenly (~n_cand / 1024 per bin).  hi bounds every candidate's distance: q~ <= thr  =>  dist <= sqrt(thr) / (1 - eps).
__global__ void k_plan_ds_fused(RefJob* job, const FusedPar* __restrict__ par, const long long* n_cand,
                                const double* cand_dist, long long cap, unsigned long long* hist, double* small,
                                unsigned* small_count, long long k, int d, int* status) {
    pdl_prologue();
    if (threadIdx.x || blockIdx.x) return;
    RefJob J;
    const bool ok = par->ok != 0 && par->thr > 0.0;
    const double hi = ok ? sqrt(par->thr) / (1.0 - par->eps) * (1.0 + 1e-6) : 1.0;
    int pk = 0;
    while ((1 << pk) < d) ++pk;
    J.src = cand_dist;
    J.src_count = reinterpret_cast<const unsigned long long*>(n_cand);
    J.src_cap = cap;
    J.transform = 0;
    J.pow2k = pk;
    J.center = 0.0;
    J.lo = pk ? 1.0 / hi : 0.0;
    J.scale = pk ? 0.0 : (double)kNB / hi;
    J.vmin = -__longlong_as_double(0x7ff0000000000000ll);
    J.kk[0] = J.kk[1] = k;
    J.nq = 1;
    J.bq[0] = J.bq[1] = -1;
    J.before = 0;
    J.hist = hist;
    J.small = small;
    J.small_count = small_count;
    J.result[0] = J.result[1] = 0.0;
    J.valid = ok ? 1 : 0;
    if (!ok) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    *job = J;
}

// the exchange block of a rank: kSmallCap gathered values followed by their count (as a double)
__global__ void k_ds_pack(const RefJob* job) {
    pdl_prologue();
    if (threadIdx.x == 0 && blockIdx.x == 0) job->small[kSmallCap] = (double)*job->small_count;
}

// blocks[world][kSmallCap + 1] (or the job's own block when world == 1): sort, read the rank -> ds
__global__ void __launch_bounds__(1024) k_ds_final(RefJob* job, const double* __restrict__ blocks, int world,
                                                   double* __restrict__ ds_out, int* status) {
    pdl_prologue();
    extern __shared__ double sm[];
    __shared__ unsigned off[65];
    RefJob& J = *job;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double* src = blocks ? blocks : J.small;
    if (threadIdx.x == 0) {
        unsigned a = 0;
        for (int r = 0; r < world; ++r) {
            off[r] = a;
            const double c = src[(size_t)r * (kSmallCap + 1) + kSmallCap];
            a += (c >= 0.0 && c <= 4.0e9) ? (unsigned)c : 0xffffffffu / 64;
        }
        off[world] = a;
    }
    __syncthreads();
    const unsigned cnt = off[world];
    const long long rk = J.kk[0] - J.before;
    const bool ok = J.valid != 0 && world <= 64 && cnt > 0 && cnt <= (unsigned)kSmallCap && rk >= 0 && rk < (long long)cnt;
    if (!ok) {
        if (threadIdx.x == 0) {
            *ds_out = nan;
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
        const double* from = src + (size_t)r * (kSmallCap + 1);
        for (int t = threadIdx.x; t < (int)c; t += blockDim.x) sm[off[r] + t] = from[t];
    }
    __syncthreads();
    bitonic_sort(sm, np2);
    if (threadIdx.x == 0) *ds_out = sm[rk];
}

// the completeness proof (header comment, item 4)
__global__ void k_fused_check(const FusedPar* __restrict__ par, const double* __restrict__ mad, int d,
                              const double* __restrict__ ds_ptr, int* status) {
    pdl_prologue();
    if (threadIdx.x || blockIdx.x) return;
    const double eps = par->eps, eta = par->eta, thr = par->thr, ds = *ds_ptr;
    bool ok = par->ok != 0 && thr > 0.0;
    for (int j = 0; j < d; ++j) {
        const double e = fabs(par->rho[j] * mad[j] - 1.0);
        ok = ok && (e <= eps);
    }
    const double lim = ds * ds * (1.0 + eps) * (1.0 + eps) * (1.0 + eta);
    ok = ok && (ds == ds) && (lim <= thr);
    if (!ok) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
}

// ---------------------------------------------------------------------------------------------------------------
struct FusedWs {
    MadWs m;
    long long capM, capU;
    FusedPar* par;
    double* est2;                  // [G2]
    unsigned long long* total;     // candidate rows marked (zeroed with the MAD counters)
    unsigned short* mask16;        // [ntiles * 16]
    int* tile_cnt;
    long long* tile_off;
    int* wprefix;                  // [nwords] candidates before every 32-row word
    unsigned long long* ealloc;    // emission slots handed out
    unsigned long long* ds_hist;   // [kNB]    } zeroed together
    unsigned* ds_small_count;      //          }
    size_t ds_zero_bytes;
    RefJob* ds_job;
    double* ds_small;              // [kSmallCap + 1]
    unsigned* fine;                // [d * kFineU32] second-stage sample counts (all-reduced as d * kFineU32 / 2 int64)
    double* est_fine;              // [d * 4] first-stage means, overwritten by the refined estimates
    size_t total_bytes;
};


inline long long round_chunk(long long x) { return (x + kFChunk - 1) / kFChunk * kFChunk; }

// entries a warp of the pass reserves at a time in a global candidate list whose expected fill is `expect`
inline int fused_chunk(long long expect) {
    long long c = expect / (160 * 16);
    c = c < 128 ? 128 : (c > 2048 ? 2048 : c);
    return (int)((c + 31) / 32 * 32);
}

// Second-stage sample: one 128-byte line out of every 2^fine lines of a column (fine = 0: off).  The first stage then
// only has to locate the windows: 8 CTAs per column instead of up to 64.
constexpr int kFineG1 = 8;
inline bool fine_valid(int fine) { return fine == 0 || (fine >= 1 && fine <= 12); }
inline long long fine_lines(long long n, int fine) { return fine ? ((n >> 4) >> fine) : 0; }
// consecutive 128-byte lines read per stratum (log2); ABCDLS_FINE_RUN overrides for experiments
inline int fine_run_shift() {
    static int rs = -1;
    if (rs < 0) {
        const char* e = getenv("ABCDLS_FINE_RUN");
        int v = e ? atoi(e) : 2;
        rs = v < 0 ? 0 : (v > 4 ? 4 : v);
    }
    return rs;
}
// how much narrower than the 64-CTA first-stage brackets the refined ones are (for the chunk reservations of the pass)
inline double fine_ratio(long long n, int fine, int world) {
    if (!fine) return 1.0;
    const double s2 = (double)fine_lines(n, fine) * 16.0 * world;
    const double r = (kZ * 0.5 / sqrt(s2 > 1.0 ? s2 : 1.0)) * 1.15 / col_delta(n, 1);
    return r > 1.0 ? 1.0 : r;
}

inline FusedWs carve_fused(void* ws, int64_t n, int d) {
    FusedWs w;
    const int G = sample_ctas(n);
    // + room for the NaN padding of the per-warp chunk reservations of the pass (one open chunk per CTA and list, and
    // what a chunk switch leaves behind)
    w.capM = round_chunk(cap_for(n, 0.03) + fused_chunk(cap_for(n, 0.03)) * 2 * 160);
    w.capU = round_chunk(cap_for(n, 0.08) + fused_chunk(cap_for(n, 0.08)) * 2 * 160);
    w.m = carve_mad(ws, d, G, w.capM, w.capU);
    char* p = reinterpret_cast<char*>(ws);
    size_t o = w.m.total;
    auto take = [&](size_t bytes) { size_t a = o; o += align_up(bytes, 256); return p ? p + a : (char*)nullptr; };
    const int trows = fused_tile_rows(d);
    const long long ntiles = (n + trows - 1) / trows;
    const long long nwords = ntiles * (trows / 32);
    const long long nblk = (nwords + 256 * kMWords - 1) / (256 * kMWords);
    w.par = (FusedPar*)take(sizeof(FusedPar));
    w.est2 = (double*)take((size_t)kFG2 * 8);
    w.total = (unsigned long long*)take(8);
    w.mask16 = (unsigned short*)take((size_t)nwords * 4 + 64);   // one bit per row of every (padded) tile
    w.tile_cnt = (int*)take((size_t)nblk * 4);
    w.tile_off = (long long*)take((size_t)nblk * 8);
    w.wprefix = (int*)take((size_t)nwords * 4);
    w.ealloc = (unsigned long long*)take(8);
    size_t z0 = o;
    w.ds_hist = (unsigned long long*)take((size_t)kNB * 8);
    w.ds_small_count = (unsigned*)take(8);
    w.ds_zero_bytes = o - z0;
    w.ds_job = (RefJob*)take(sizeof(RefJob));
    w.ds_small = (double*)take((size_t)(kSmallCap + 1) * 8);
    w.fine = (unsigned*)take((size_t)d * kFineU32 * 4);
    w.est_fine = (double*)take((size_t)d * 4 * 8);
    w.total_bytes = o;
    return w;
}

}  // namespace

extern "C" {

size_t abcdls_fused_workspace_bytes(int64_t n, int d) { return carve_fused(nullptr, n, d).total_bytes; }

// 1 if the one-pass path can stream this table (TMA needs 16-byte aligned columns; d <= 16 fits one ring stage)
int abcdls_fused_supported(const double* ss, int64_t n, int d, int64_t ld) {
    if (!(d >= 1 && d <= kFMaxD && n >= 4 * kSample && n < (1ll << 31) && ld >= n && (ld & 1) == 0 &&
          (reinterpret_cast<uintptr_t>(ss) & 15) == 0))
        return 0;
    return encode_tiled_fn() ? 1 : 0;
}

inline int fused_G(long long n, int world, int fine) {
    const int g = sample_ctas_w(n, world);
    return (fine && g > kFineG1) ? kFineG1 : g;
}

#define FINE_ARG 0   /* entry points that take `fine` redefine it around their prologue */
#define FUSED_PROLOGUE(name)                                                                                        \
    ABCDLS_CHECK_ARG(h, h && ss && ws && abcdls_fused_supported(ss, n, d, ld) && world >= 1 && world <= 64, name);   \
    FusedWs w = carve_fused(ws, n, d);                                                                              \
    if (ws_bytes < w.total_bytes) {                                                                                 \
        h->err = "fused workspace too small";                                                                       \
        return ABCDLS_E_WORKSPACE;                                                                                  \
    }                                                                                                               \
    cudaStream_t st = (cudaStream_t)stream;                                                                         \
    const int G = fused_G(n, world, FINE_ARG);  /* CTAs per column of the first-stage sampling kernel */            \
    const int G2 = (world > 1 && kFG2 / world < 8) ? 8 : kFG2 / (world > 1 ? world : 1);                            \
    (void)G;                                                                                                        \
    (void)G2

// Stage 1: column samples.  est_out: doubles a multi-rank caller all-reduces (sum) before stage 2.
#undef FINE_ARG
#define FINE_ARG fine
int abcdls_fused_sample_cols(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int fine,
                             void* ws, size_t ws_bytes, double** est_out, size_t* est_count, void* stream) {
    FUSED_PROLOGUE("fused_sample_cols");
    ABCDLS_CHECK_ARG(h, fine_valid(fine), "fused_sample_cols: fine");
    double delta = kZ * 0.5 / sqrt((double)G * kSample * world);
    if (delta < 2.0 / kSample) delta = 2.0 / kSample;
    if (fine) ABCDLS_CUDA(h, cudaMemsetAsync(w.fine, 0, (size_t)d * kFineU32 * 4, st));
    const size_t smem = (size_t)2 * kSample * sizeof(double);
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_sample_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.m.cnt, 0, w.m.zero_bytes, st));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.total, 0, 8, st));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.ealloc, 0, 8, st));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.ds_hist, 0, w.ds_zero_bytes, st));
    launch_pdl(k_sample_cols, dim3(dim3(G, d)), dim3(512), smem, st, ss, n, ld, 0x5EEDu, delta, w.m.est);
    ABCDLS_LAUNCH_CHECK(h);
    if (est_out) *est_out = w.m.est;
    if (est_count) *est_count = (size_t)G * d * 4;
    return ABCDLS_OK;
}

// Stage 1b (fine > 0 only): first-stage windows from the (reduced) column samples, then the second-stage COUNTING
// sample (one 128-byte line out of every 2^fine of each column).  cnt_out: int64 words to all-reduce (sum) before
// stage 2 (pairs of 32-bit counts: the sums stay below 2^32).
int abcdls_fused_sample_fine(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int fine,
                             void* ws, size_t ws_bytes, int64_t** cnt_out, size_t* cnt_count, void* stream) {
    FUSED_PROLOGUE("fused_sample_fine");
    ABCDLS_CHECK_ARG(h, fine >= 1 && fine_valid(fine), "fused_sample_fine: fine");
    launch_pdl(k_make_brackets, dim3(d), dim3(32), 0, st, w.m.est, G, d, 1.0 / world, w.m.brk, w.est_fine);
    const int rs = fine_run_shift();
    const int rq = rs > 2 ? rs : 2;
    const long long L = fine_lines(n, fine) >> rq << rq;    // whole runs, and four lines per half-warp step
    if (L > 0) {
        long long gx = (L + 64 * 4 - 1) / (64 * 4);
        long long mx = (long long)h->sm_count * 2 / d;   // 63 registers x 512 threads: two CTAs per SM, one wave
        if (mx < 1) mx = 1;
        if (gx > mx) gx = mx;
        launch_pdl(k_sample_fine, dim3(dim3((unsigned)gx, d)), dim3(512), 0, st, ss, n, ld, w.m.brk, L, fine, rs, 0xF17Eu, w.fine);
    }
    ABCDLS_LAUNCH_CHECK(h);
    if (cnt_out) *cnt_out = reinterpret_cast<int64_t*>(w.fine);
    if (cnt_count) *cnt_count = (size_t)d * kFineU32 / 2;
    return ABCDLS_OK;
}

// Stage 2: brackets + scale estimates from the (reduced) column samples, then the row sample of approximate
// distances.  est_out: doubles to all-reduce (sum) before stage 3.
int abcdls_fused_sample_dist(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* target,
                             int64_t nacc, int64_t n_global, int world, int fine, void* ws, size_t ws_bytes,
                             double** est_out, size_t* est_count, void* stream) {
    FUSED_PROLOGUE("fused_sample_dist");
    if (!(target && nacc >= 1 && n_global >= nacc && fine_valid(fine))) {
        h->err = "invalid argument: fused_sample_dist: target " + std::string(target ? "ok" : "NULL") + ", nacc " +
                 std::to_string((long long)nacc) + ", n_global " + std::to_string((long long)n_global) + ", fine " +
                 std::to_string(fine);
        return ABCDLS_E_INVALID;
    }
    if (fine) {
        launch_pdl(k_fine_brackets, dim3(d), dim3(256), 0, st, w.fine, w.m.brk, kZ, w.est_fine);
        launch_pdl(k_make_brackets, dim3(d), dim3(32), 0, st, w.est_fine, 1, d, 1.0, w.m.brk, nullptr);
        launch_pdl(k_make_fused_scale, dim3(1), dim3(32), 0, st, w.est_fine, 1, d, 1.0, target, w.par);
    } else {
        launch_pdl(k_make_brackets, dim3(d), dim3(32), 0, st, w.m.est, G, d, 1.0 / world, w.m.brk, nullptr);
        launch_pdl(k_make_fused_scale, dim3(1), dim3(32), 0, st, w.m.est, G, d, 1.0 / world, target, w.par);
    }
    const double p = (double)nacc / (double)n_global;
    const double delta = kZ * sqrt(p * (1.0 - p) / ((double)G2 * kSampleQ * world));
    launch_pdl(k_sample_qdist, dim3(G2), dim3(512), (size_t)kSampleQ * sizeof(double), st, ss, n, d, ld, w.par, 0xD157u, p, delta, w.est2);
    ABCDLS_LAUNCH_CHECK(h);
    if (est_out) *est_out = w.est2;
    if (est_count) *est_count = (size_t)G2;
    return ABCDLS_OK;
}

// Stage 3: THE table pass.  counts_out: int64 block (counts | band counts | median histogram) to all-reduce (sum).
int64_t abcdls_fused_emit_cap(int64_t cap) { return cap + (int64_t)160 * kFWarps * kFEChunk; }
// doubles per emitted candidate record (the row pitch of cand_ss): 16 for d <= 16, 32 for d <= 32
int abcdls_fused_emit_pitch(int d) { return fused_emit_pitch(d); }

int abcdls_fused_pass(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int fine,
                      double* cand_ss, int32_t* cand_lrow, int64_t ecap, void* ws, size_t ws_bytes, int* status_dev,
                      int64_t** counts_out, size_t* counts_count, void* stream) {
    FUSED_PROLOGUE("fused_pass");
    ABCDLS_CHECK_ARG(h, status_dev && cand_ss && cand_lrow && ecap >= 1 && ecap < (1ll << 31) && fine_valid(fine),
                     "fused_pass: args");
    const double narrow = fine_ratio(n, fine, world);   // the refined brackets keep that much less per list
    ABCDLS_CUDA(h, cudaMemsetAsync(cand_lrow, 0xff, (size_t)ecap * 4, st));   // -1 = unused slot
    launch_pdl(k_make_fused_thr, dim3(1), dim3(32), 0, st, w.est2, G2, 1.0 / world, d, w.par);
    const int nstages = fused_stages(d);
    const int trows = fused_tile_rows(d);
    const bool wide = d > kFNarrowD;
    const size_t smem = kFHeader + (size_t)nstages * d * trows * 8 + kFQBytes;
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_fused_pass, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kFHeader + kFRingBytes + kFQBytes));
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_fused_pass_wide, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kFHeader + kFRingBytes + kFQBytes));
    const long long ntiles = (n + trows - 1) / trows;
    long long grid = h->sm_count;
    if (grid > ntiles) grid = ntiles;
    const long long tpc = (ntiles + grid - 1) / grid;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        h->err = "cuTensorMapEncodeTiled is not available from this driver";
        return ABCDLS_E_CUDA;
    }
    alignas(64) CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)d};
        const cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
        const cuuint32_t box[2] = {(cuuint32_t)trows, (cuuint32_t)d};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ss), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
        if (r != CUDA_SUCCESS) {
            h->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
            return ABCDLS_E_CUDA;
        }
    }
    const int chM = fused_chunk((long long)(cap_for(n, 0.03) * narrow)), chU = fused_chunk((long long)(cap_for(n, 0.08) * narrow));
    if (wide)
        launch_pdl(k_fused_pass_wide, dim3((unsigned)grid), dim3(kFThreads), smem, st, tmap, n, ntiles, d, nstages, w.m.brk,
                   w.par, w.m.counts, w.m.candM, w.capM, w.m.candU, w.capU, w.m.cnt,
                   reinterpret_cast<unsigned char*>(w.mask16), cand_ss, ecap, cand_lrow, w.ealloc, tpc, chM, chU,
                   status_dev);
    else
        launch_pdl(k_fused_pass, dim3((unsigned)grid), dim3(kFThreads), smem, st, tmap, n, ntiles, d, nstages, w.m.brk,
                   w.par, w.m.counts, w.m.candM, w.capM, w.m.candU, w.capU, w.m.cnt, w.mask16, cand_ss, ecap, cand_lrow,
                   w.ealloc, tpc, chM, chU, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    {
        long long gx = (w.capM + 256 * 8 - 1) / (256 * 8);
        long long mx = (long long)h->sm_count * 8 / d;
        if (mx < 1) mx = 1;
        if (gx > mx) gx = mx;
        launch_pdl(k_fused_histM, dim3(dim3((unsigned)gx, d)), dim3(256), 0, st, w.m.candM, w.capM, w.m.cnt, w.m.brk, w.m.histM, w.m.counts);
        launch_pdl(k_copy_u64, dim3(1), dim3(128), 0, st, w.m.gcnt, w.m.cnt, 2 * d);
        ABCDLS_LAUNCH_CHECK(h);
    }
    if (counts_out) *counts_out = reinterpret_cast<int64_t*>(w.m.counts);
    if (counts_count) *counts_count = w.m.redA_count;
    return ABCDLS_OK;
}

#undef FINE_ARG
#define FINE_ARG 0
// Stage 4: median / MAD refinement, phases 0..3 with the same exchanges as abcdls_mad_fast_refine.
int abcdls_fused_refine(abcdls_handle_t h, const double* ss, int64_t n, int64_t n_global, int d, int64_t ld, int phase,
                        int world, const double* gath_small, const int32_t* gath_cnt, void* ws, size_t ws_bytes,
                        double* med, double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1,
                        size_t* x1_count, void* stream) {
    FUSED_PROLOGUE("fused_refine");
    ABCDLS_CHECK_ARG(h, med && mad && status_dev && phase >= 0 && phase <= 3, "fused_refine: args");
    return refine_impl(h, w.m, w.capM, w.capU, false, n_global, d, phase, world, gath_small, gath_cnt, med, mad,
                       status_dev, x0, x0_count, x1, x1_count, st);
}

// Stage 5a (needs only the pass): marked rows -> ascending global candidate rows (cand_row), n_cand (device int64).
// Independent of the median / MAD refinement, so a caller can start work that only needs the candidate rows
// (e.g. pre-gathering their parameters on a second stream) while stage 4 runs.
int abcdls_fused_rows(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int64_t row_offset,
                      int64_t cap, int64_t* cand_row, int64_t* n_cand, void* ws, size_t ws_bytes, int* status_dev,
                      void* stream) {
    FUSED_PROLOGUE("fused_rows");
    ABCDLS_CHECK_ARG(h, cand_row && n_cand && status_dev && cap >= 1, "fused_rows");
    const int trows = fused_tile_rows(d);
    const long long ntiles = (n + trows - 1) / trows;
    const long long nwords = ntiles * (trows / 32);
    const long long nblk = (nwords + 256 * kMWords - 1) / (256 * kMWords);
    const unsigned* bitmap = reinterpret_cast<const unsigned*>(w.mask16);
    launch_pdl(k_mask_count, dim3((unsigned)nblk), dim3(256), 0, st, bitmap, nwords, w.tile_cnt, w.total);
    launch_pdl(k_scan_tiles, dim3(1), dim3(1024), 0, st, w.tile_cnt, nblk, w.tile_off);
    launch_pdl(k_mask_write, dim3((unsigned)nblk), dim3(256), 0, st, bitmap, nwords, w.tile_off, row_offset, cap, (long long*)cand_row,
                                                 w.wprefix, w.total, (long long*)n_cand, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// Stage 5b (needs the exact MADs): R's exact distance of every candidate from the block the pass emitted
// (cand_ss / cand_lrow / ecap as given to the pass), scattered to its position in the ascending candidate list;
// perm[i] = slot of candidate i in that block.
int abcdls_fused_exact(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, const double* mad,
                       const double* target, int64_t cap, double* cand_dist, const double* cand_ss,
                       const int32_t* cand_lrow, int64_t ecap, int64_t* perm, void* ws, size_t ws_bytes, int* status_dev,
                       void* stream) {
    FUSED_PROLOGUE("fused_exact");
    ABCDLS_CHECK_ARG(h, mad && target && cand_dist && cand_ss && cand_lrow && perm && status_dev && cap >= 1 && ecap >= 1,
                     "fused_exact");
    const unsigned* bitmap = reinterpret_cast<const unsigned*>(w.mask16);
    long long g = (ecap + 255) / 256;
    if (g > (long long)h->sm_count * 16) g = (long long)h->sm_count * 16;
    launch_pdl(k_cand_exact, dim3((unsigned)g), dim3(256), 0, st, cand_ss, ecap, cand_lrow, w.ealloc, d, fused_emit_pitch(d), mad, target, bitmap, w.wprefix, cap,
                                              cand_dist, (long long*)perm, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// Stage 5 = 5a + 5b
int abcdls_fused_candidates(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world,
                            const double* mad, const double* target, int64_t row_offset, int64_t cap, int64_t* cand_row,
                            double* cand_dist, const double* cand_ss, const int32_t* cand_lrow, int64_t ecap,
                            int64_t* perm, int64_t* n_cand, void* ws, size_t ws_bytes, int* status_dev, void* stream) {
    int rc = abcdls_fused_rows(h, ss, n, d, ld, world, row_offset, cap, cand_row, n_cand, ws, ws_bytes, status_dev, stream);
    if (rc) return rc;
    return abcdls_fused_exact(h, ss, n, d, ld, world, mad, target, cap, cand_dist, cand_ss, cand_lrow, ecap, perm, ws,
                              ws_bytes, status_dev, stream);
}

// Stage 5c: exact ds among the candidates of all ranks.
//   phase 0: plan + histogram of this rank's candidate distances   -> exchange: all-reduce (sum) x0 (kNB int64)
//   phase 1: locate the bin + gather its values                     -> exchange: all-gather x0 (8193 doubles: values, count)
//   phase 2: sort the gathered values -> ds_out (gath = the gathered blocks, NULL on a single rank)
int abcdls_fused_ds(abcdls_handle_t h, int64_t n, int d, int phase, int world, const double* gath,
                    const double* cand_dist, const int64_t* n_cand, int64_t cap, int64_t nacc, double* ds_out, void* ws,
                    size_t ws_bytes, int* status_dev, void** x0, size_t* x0_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && cand_dist && n_cand && ds_out && status_dev && d >= 1 && d <= kFMaxD && phase >= 0 &&
                            phase <= 2 && world >= 1 && world <= 64 && nacc >= 1 && cap >= 1,
                     "fused_ds");
    FusedWs w = carve_fused(ws, n, d);
    if (ws_bytes < w.total_bytes) {
        h->err = "fused workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long gg = (cap + 2047) / 2048;
    if (gg > (long long)h->sm_count * 4) gg = (long long)h->sm_count * 4;
    if (x0) *x0 = nullptr;
    if (phase == 0) {
        launch_pdl(k_plan_ds_fused, dim3(1), dim3(32), 0, st, w.ds_job, w.par, (const long long*)n_cand, cand_dist, cap, w.ds_hist, w.ds_small,
                                          w.ds_small_count, nacc - 1, d, status_dev);
        launch_pdl(k_refine_hist, dim3(dim3((unsigned)gg, 1)), dim3(256), 0, st, w.ds_job, cap);
        if (x0) { *x0 = w.ds_hist; *x0_count = (size_t)kNB; }
    } else if (phase == 1) {
        launch_pdl(k_refine_pick, dim3(1), dim3(256), 0, st, w.ds_job, status_dev);
        launch_pdl(k_refine_gather, dim3(dim3((unsigned)gg, 1)), dim3(256), 0, st, w.ds_job, cap);
        launch_pdl(k_ds_pack, dim3(1), dim3(32), 0, st, w.ds_job);
        if (x0) { *x0 = w.ds_small; *x0_count = (size_t)kSmallCap + 1; }
    } else {
        const size_t smem = (size_t)kSmallCap * sizeof(double);
        ABCDLS_ONCE(h, cudaFuncSetAttribute(k_ds_final, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_pdl(k_ds_final, dim3(1), dim3(1024), smem, st, w.ds_job, gath, gath ? world : 1, ds_out, status_dev);
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// Stage 6 (after the exact ds has been selected among the candidates of all ranks): completeness proof.
int abcdls_fused_check(abcdls_handle_t h, int64_t n, int d, const double* mad, const double* ds, void* ws,
                       size_t ws_bytes, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && mad && ds && status_dev && d >= 1 && d <= kFMaxD, "fused_check");
    FusedWs w = carve_fused(ws, n, d);
    if (ws_bytes < w.total_bytes) {
        h->err = "fused workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    launch_pdl(k_fused_check, dim3(1), dim3(32), 0, (cudaStream_t)stream, w.par, mad, d, ds, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
