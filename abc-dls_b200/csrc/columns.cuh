// columns.cuh — ABC-DLS's per-column mode as ONE pass: P independent 1-D abc(method = "rejection") problems that share
// the N rows of an N x P statistics table (reference: src/Classes/ABC.py:2805-2840 abc_params_min_max - the SMC round;
// 1949-1953 abc_params).  Included at the end of bracket.cu: shares its column samples, brackets, band refinement and
// workspace.
//
// Per column j R needs three order statistics: median(x), median|x - median| (-> MAD M_j) and the nacc-th smallest
// dist = |x / M_j - t_j / M_j|.  For ONE statistic the distance is a monotone function of |x - t_j| on either side of
// the target (IEEE division by a positive constant, subtraction, square and square root are all monotone), so the
// tolerance cut is a third BAND around the target whose width does not depend on the MAD:
//     T band:  fl|x - t_j| <= R_j,   R_j = an upper bracket of the nacc-th smallest |x - t_j| from the column sample.
// The pass therefore classifies every element against the median band, the MAD band (exactly as k_bracket_pass) and
// the T band, whose members (value, local row) are kept.  Afterwards: exact medians / MADs (shared refinement), R's
// exact distance for the ~1.1 nacc T-band members of every column, ds_j by the radix select over all ranks, the proof
// that nothing outside the band can have dist <= ds_j (the distance at the band edges is > ds_j; else
// ABCDLS_S_FASTPATH_MISS and the caller falls back to P separate calls), R's cap rule (ties at ds_j enter in row
// order, lower ranks first), and min / max / sum of the accepted parameters.
//   sample -> [all-reduce est] -> pass -> [all-reduce counts] -> refine phases 0..3 (as abcdls_mad_fast_refine)
//   -> exact -> [select: abcdls_select_*] -> count -> [all-gather n_lt, n_eq] -> row keys -> [local select] -> accept
//   -> final [-> in-kernel exchange]
#include "peer.cuh"

namespace {

constexpr int kCStageT = 288;   // staged T-band members per warp (a whole warp-tile of 256 always fits after a flush)

struct ColsWs {
    MadWs m;                        // est has room for 5 estimates per (CTA, column): 4 of the column brackets | T
    double* RT;                     // [P] T-band radius
    unsigned long long* cntT;       // [P] T-band members appended by THIS rank
    unsigned long long* red;        // [P] } members over all ranks   (reduce block together with m.counts .. m.histM)
    unsigned long long* cnt2;       // [P][2] n_lt, n_eq of this rank
    double* keys;                   // [P][capT] row keys of the cap rule
    double* valT;                   // [P][capT]
    int* rowT;                      // [P][capT]
    double* distT;                  // [P][capT]
    double* part;                   // [P][gx][6]  block partials: min, max, sum, count of the accepted parameters, min / max of x
    long long capT;
    int gx_acc;
    size_t total;
};

inline long long cols_capT(long long n, long long nacc_local_hint) {
    // the band holds ~(tol + 5.5 sigma) of the rows; sized for twice the expected share + slack, never more than n
    long long c = 2 * nacc_local_hint + (long long)(0.004 * (double)n) + 65536;
    return c > n ? n : c;
}

inline ColsWs carve_cols(void* ws, int P, long long n, long long capT, int sm_count) {
    ColsWs w;
    const int G = sample_ctas(n);
    const int Geff = (G * 5 + 3) / 4;
    w.m = carve_mad(ws, P, Geff, cap_for(n, 0.03), cap_for(n, 0.08));
    char* p = reinterpret_cast<char*>(ws);
    size_t o = w.m.total;
    auto take = [&](size_t bytes) { size_t a = o; o += align_up(bytes, 256); return p ? p + a : (char*)nullptr; };
    w.capT = capT;
    w.RT = (double*)take((size_t)P * 8);
    w.cntT = (unsigned long long*)take((size_t)P * 8);
    w.red = (unsigned long long*)take((size_t)P * 8);
    w.cnt2 = (unsigned long long*)take((size_t)P * 2 * 8);
    w.valT = (double*)take((size_t)P * capT * 8);
    w.rowT = (int*)take((size_t)P * capT * 4);
    w.distT = (double*)take((size_t)P * capT * 8);
    w.keys = (double*)take((size_t)P * capT * 8);
    w.gx_acc = sm_count * 4 / P < 1 ? 1 : sm_count * 4 / P;
    w.part = (double*)take((size_t)P * w.gx_acc * 6 * 8);
    w.total = o;
    return w;
}

// K1: per-column sample -> (m_lo, m_hi, r_lo, r_hi) as k_sample_cols, plus the (p + deltaT) quantile of |x - t_j|.
// est[(g * P + j) * 4 + q] for the four bracket estimates, estT[g * P + j] behind them.  grid (G, P), 512 threads
__global__ void __launch_bounds__(512) k_sample_cols_t(const double* __restrict__ ss, long long n, long long ld,
                                                       unsigned seed, double delta, const double* __restrict__ target,
                                                       double pT, double* __restrict__ est, double* __restrict__ estT) {
    extern __shared__ double sm[];
    double* s = sm;            // kSample
    double* a = sm + kSample;  // kSample
    const int j = blockIdx.y;
    const double* col = ss + (size_t)j * ld;
    for (int t = threadIdx.x; t < kSample; t += blockDim.x) s[t] = col[sample_row(seed + 977u * j, blockIdx.x, t, n)];
    __syncthreads();
    bitonic_sort(s, kSample);
    const double mhat = s[kSample / 2];
    for (int t = threadIdx.x; t < kSample; t += blockDim.x) a[t] = fabs(s[t] - mhat);
    __syncthreads();
    bitonic_merge_up(a, kSample);     // |s - mhat| of a sorted sample is bitonic: one merge sorts it
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
    __syncthreads();
    const double tj = target[j];
    for (int t = threadIdx.x; t < kSample; t += blockDim.x) a[t] = fabs(s[t] - tj);
    __syncthreads();
    // |s - t| falls and rises again too (or is monotone when the target lies outside the sample): still bitonic
    bitonic_merge_up(a, kSample);
    if (threadIdx.x == 0) {
        int it = (int)ceil(pT * kSample);
        it = max(1, min(kSample - 1, it));
        estT[(size_t)blockIdx.x * gridDim.y + j] = a[it];
    }
}

__global__ void k_make_tband(const double* __restrict__ estT, int G, int P, double inv_world, double* __restrict__ RT) {
    const int j = blockIdx.x, lane = threadIdx.x;
    double m = 0.0;
    for (int g = lane; g < G; g += 32) m += estT[(size_t)g * P + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_down_sync(0xffffffffu, m, o);
    if (lane == 0) RT[j] = m / G * inv_world;
}

__device__ __noinline__ void stage_flush_t(const double* sv, const int* sr, int fill, double* __restrict__ gval,
                                           int* __restrict__ grow, long long gcap, unsigned long long* gcnt) {
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0ull;
    if (lane == 0) base = atomicAdd(gcnt, (unsigned long long)fill);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int t = lane; t < fill; t += 32)
        if ((long long)base + t < gcap) {
            gval[base + t] = sv[t];
            grow[base + t] = sr[t];
        }
    __syncwarp();
}

// K2: the pass.  k_bracket_pass (same loop, same counters, same M / U staging) + the T band.  grid (gx, P)
__global__ void __launch_bounds__(kAThreads, 3) k_columns_pass(const double* __restrict__ ss, long long n, long long ld,
                                                               const ColBrk* __restrict__ brk,
                                                               const double* __restrict__ target,
                                                               const double* __restrict__ RTv,
                                                               unsigned long long* __restrict__ counts,
                                                               unsigned long long* __restrict__ histM,
                                                               double* __restrict__ candM, long long capM,
                                                               double* __restrict__ candU, long long capU,
                                                               unsigned long long* __restrict__ cnt,
                                                               double* __restrict__ valT, int* __restrict__ rowT,
                                                               long long capT, unsigned long long* __restrict__ cntT) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int NW = kAThreads / 32;
    unsigned* hist = reinterpret_cast<unsigned*>(smraw);                       // [kNB]
    double* stM = reinterpret_cast<double*>(smraw + kNB * 4);                  // [NW][kStageM]
    double* stU = stM + NW * kStageM;                                          // [NW][kStageU]
    double* stTv = stU + NW * kStageU;                                         // [NW][kCStageT]
    int* stTr = reinterpret_cast<int*>(stTv + NW * kCStageT);                  // [NW][kCStageT]
    __shared__ unsigned tot2[2];
    const int j = blockIdx.y;
    const double* __restrict__ col = ss + (size_t)j * ld;
    const ColBrk B = brk[j];
    const double c = B.c, r0 = B.r0, r1 = B.r1, r2 = B.r2, nr0 = -B.r0, loM = B.loM, scaleM = B.scaleM;
    const double tj = target[j], RT = RTv[j];
    for (int t = threadIdx.x; t < kNB; t += kAThreads) hist[t] = 0u;
    if (threadIdx.x < 2) tot2[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long warp_id = (long long)blockIdx.x * NW + wid;
    const long long warp_stride = (long long)gridDim.x * NW;
    const bool vec = ((reinterpret_cast<uintptr_t>(col) & 15) == 0);
    const long long nfull = n / kAWarpTile;
    const long long ntiles = (n + kAWarpTile - 1) / kAWarpTile;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned c_lt = 0, c_in = 0;
    int fillM = 0, fillU = 0, fillT = 0;
    double* myM = stM + wid * kStageM;
    double* myU = stU + wid * kStageU;
    double* myTv = stTv + wid * kCStageT;
    int* myTr = stTr + wid * kCStageT;
    double* gM = candM + (size_t)j * capM;
    double* gU = candU + (size_t)j * capU;
    double* gTv = valT + (size_t)j * capT;
    int* gTr = rowT + (size_t)j * capT;
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
    auto process = [&](long long wt, const double (&x)[kAItems]) {
        if (fillM > kStageM - kAWarpTile) {
            stage_flush<true>(myM, fillM, gM, capM, cntM, hist, loM, scaleM);
            fillM = 0;
        }
        if (fillU > kStageU - kAWarpTile) {
            stage_flush<false>(myU, fillU, gU, capU, cntU, nullptr, 0.0, 0.0);
            fillU = 0;
        }
        if (fillT > kCStageT - kAWarpTile) {
            stage_flush_t(myTv, myTr, fillT, gTv, gTr, capT, &cntT[j]);
            fillT = 0;
        }
        const int row0 = (int)(wt * kAWarpTile) + 2 * lane;
#pragma unroll
        for (int u = 0; u < kAItems; ++u) {
            const double sgn = __dsub_rn(x[u], c);
            const double y = fabs(sgn);
            const bool in1 = y < r1;
            const bool inM = y <= r0;
            const bool inU = !in1 && y <= r2;
            const bool inT = fabs(__dsub_rn(x[u], tj)) <= RT;          // NaN padding of the last tile: false
            if (sgn < nr0) ++c_lt;
            if (in1) ++c_in;
            const unsigned bM = __ballot_sync(0xffffffffu, inM);
            const unsigned bU = __ballot_sync(0xffffffffu, inU);
            const unsigned bT = __ballot_sync(0xffffffffu, inT);
            if (bM) {
                if (inM) myM[fillM + __popc(bM & lt_mask)] = x[u];
                fillM += __popc(bM);
            }
            if (bU) {
                if (inU) myU[fillU + __popc(bU & lt_mask)] = x[u];
                fillU += __popc(bU);
            }
            if (bT) {
                if (inT) {
                    const int q = fillT + __popc(bT & lt_mask);
                    myTv[q] = x[u];
                    myTr[q] = row0 + (u >> 1) * 64 + (u & 1);
                }
                fillT += __popc(bT);
            }
        }
    };

    double xa[kAItems], xb[kAItems];
    long long wt = warp_id;
    if (wt < ntiles) load_tile(wt, xa);
    while (wt < ntiles) {
        long long wnext = wt + warp_stride;
        if (wnext < ntiles) load_tile(wnext, xb);
        process(wt, xa);
        wt = wnext;
        if (wt >= ntiles) break;
        wnext = wt + warp_stride;
        if (wnext < ntiles) load_tile(wnext, xa);
        process(wt, xb);
        wt = wnext;
    }
    __syncwarp();
    if (fillM) stage_flush<true>(myM, fillM, gM, capM, cntM, hist, loM, scaleM);
    if (fillU) stage_flush<false>(myU, fillU, gU, capU, cntU, nullptr, 0.0, 0.0);
    if (fillT) stage_flush_t(myTv, myTr, fillT, gTv, gTr, capT, &cntT[j]);
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

// R's distance for ONE statistic: sqrt((x / M - t / M)^2), every operation rounded like R does it
__device__ __forceinline__ double dist_1d(double x, const double2 dr, double tsc) {
    const double a = __dsub_rn(scaled_value(x, dr), tsc);
    return __dsqrt_rn(__dmul_rn(a, a));
}
__device__ __forceinline__ void scale_1d(double m, double t, double2& dr, double& tsc) {
    const double dv = (m == 0.0) ? 1.0 : m;
    const double av = fabs(dv);
    dr = make_double2(dv, (av > 1e-90 && av < 1e90) ? 1.0 / dv : 0.0);
    tsc = (m == 0.0) ? t : t / dv;
}

// K3: exact distances of the T-band members.  grid (gx, P)
__global__ void __launch_bounds__(256) k_cols_exact(const double* __restrict__ valT, long long capT,
                                                    const unsigned long long* __restrict__ cntT,
                                                    const double* __restrict__ mad, const double* __restrict__ target,
                                                    double* __restrict__ distT, long long* __restrict__ cnt_out,
                                                    int* status) {
    const int j = blockIdx.y;
    double2 dr;
    double tsc;
    scale_1d(mad[j], target[j], dr, tsc);
    long long nc = (long long)cntT[j];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (nc > capT) atomicOr(status, ABCDLS_S_FASTPATH_MISS);     // members were lost
        cnt_out[j] = nc > capT ? capT : nc;
    }
    if (nc > capT) nc = capT;
    bool nonfinite = false;
    const double* v = valT + (size_t)j * capT;
    double* o = distT + (size_t)j * capT;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride) {
        const double dd = dist_1d(v[i], dr, tsc);
        nonfinite |= !isfinite(dd);
        o[i] = dd;
    }
    if (__any_sync(0xffffffffu, nonfinite) && (threadIdx.x & 31) == 0) atomicOr(status, ABCDLS_S_NONFINITE);
}

// K4: n_lt = #(dist < ds), n_eq = #(dist == ds) of this rank's members, and (block 0 of a column) the completeness
// proof: enough members over all ranks, and the distance at both edges of the band is > ds.  grid (gx, P)
__global__ void __launch_bounds__(256) k_cols_count(const double* __restrict__ distT, long long capT,
                                                    const long long* __restrict__ cnt, const double* __restrict__ sel,
                                                    const double* __restrict__ mad, const double* __restrict__ target,
                                                    const double* __restrict__ RTv,
                                                    const unsigned long long* __restrict__ gcntT, long long nacc,
                                                    unsigned long long* __restrict__ cnt2, int* status) {
    const int j = blockIdx.y;
    const double ds = sel[2 * j];
    const long long nc = cnt[j];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double2 dr;
        double tsc;
        scale_1d(mad[j], target[j], dr, tsc);
        const double t = target[j], R = RTv[j] * (1.0 - 1e-12);
        const double e0 = dist_1d(t + R, dr, tsc), e1 = dist_1d(t - R, dr, tsc);
        if ((long long)gcntT[j] < nacc || !(e0 > ds) || !(e1 > ds)) atomicOr(status, ABCDLS_S_FASTPATH_MISS);
    }
    const double* v = distT + (size_t)j * capT;
    unsigned lt = 0, eq = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride) {
        const double dd = v[i];
        lt += dd < ds;
        eq += dd == ds;
    }
    lt = warp_sum(lt);
    eq = warp_sum(eq);
    if ((threadIdx.x & 31) == 0) {
        if (lt) atomicAdd(&cnt2[2 * j], (unsigned long long)lt);
        if (eq) atomicAdd(&cnt2[2 * j + 1], (unsigned long long)eq);
    }
}

// K5: R's cap rule - wt1 <- dist <= ds; wt1 & cumsum(wt1) <= nacc - keeps the FIRST nacc rows, in row order, of the
// whole set {dist <= ds} (not "everything below ds plus some ties").  The members are unordered, so the cut is a row
// number: keys[j][i] = local row of member i if dist <= ds, +inf otherwise; the caller selects this rank's quota-th
// smallest key per column (abcdls_select_* without the cross-rank reduction).  grid (gx, P)
__global__ void __launch_bounds__(256) k_cols_rowkeys(const double* __restrict__ distT, const int* __restrict__ rowT,
                                                      long long capT, const long long* __restrict__ cnt,
                                                      const double* __restrict__ sel, double* __restrict__ keys) {
    const int j = blockIdx.y;
    const double ds = sel[2 * j];
    const long long nc = cnt[j];
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride)
        keys[(size_t)j * capT + i] = distT[(size_t)j * capT + i] <= ds ? (double)rowT[(size_t)j * capT + i] : inf;
}

// K6: accepted members -> parameter value (column-major or row-major table, device or pinned host) -> block partials
// {min, max, sum, count}; optionally the accepted (local row, value, distance) lists (unordered).  grid (gx, P)
__global__ void __launch_bounds__(256) k_cols_accept(const double* __restrict__ distT, const int* __restrict__ rowT,
                                                     const double* __restrict__ valT,
                                                     long long capT, const long long* __restrict__ cnt,
                                                     const double* __restrict__ sel,
                                                     const double* __restrict__ thr_row,
                                                     const double* __restrict__ par, long long par_cs, long long par_rs,
                                                     double* __restrict__ part, long long acc_cap,
                                                     int* __restrict__ acc_row, double* __restrict__ acc_val,
                                                     double* __restrict__ acc_dist,
                                                     unsigned long long* __restrict__ acc_cnt) {
    __shared__ double sh[6][8];
    const int j = blockIdx.y;
    const double ds = sel[2 * j];
    const long long nc = cnt[j];
    const double thr = thr_row[2 * j];            // this rank keeps the members with dist <= ds and local row <= thr
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double mn = inf, mx = -inf, sum = 0.0, num = 0.0, xmn = inf, xmx = -inf;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += stride) {
        const double dd = distT[(size_t)j * capT + i];
        const int row = rowT[(size_t)j * capT + i];
        const bool acc = dd <= ds && (double)row <= thr;
        if (acc) {
            const double y = par[(size_t)j * par_cs + (size_t)row * par_rs];
            const double xv = valT[(size_t)j * capT + i];
            xmn = fmin(xmn, xv);
            xmx = fmax(xmx, xv);
            mn = fmin(mn, y);
            mx = fmax(mx, y);
            sum += y;
            num += 1.0;
            if (acc_row) {
                const unsigned long long p = atomicAdd(&acc_cnt[j], 1ull);
                if ((long long)p < acc_cap) {
                    acc_row[(size_t)j * acc_cap + p] = row;
                    acc_val[(size_t)j * acc_cap + p] = y;
                    acc_dist[(size_t)j * acc_cap + p] = dd;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        num += __shfl_xor_sync(0xffffffffu, num, o);
        xmn = fmin(xmn, __shfl_xor_sync(0xffffffffu, xmn, o));
        xmx = fmax(xmx, __shfl_xor_sync(0xffffffffu, xmx, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sh[0][wid] = mn;
        sh[1][wid] = mx;
        sh[2][wid] = sum;
        sh[3][wid] = num;
        sh[4][wid] = xmn;
        sh[5][wid] = xmx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            sh[0][0] = fmin(sh[0][0], sh[0][w]);
            sh[1][0] = fmax(sh[1][0], sh[1][w]);
            sh[2][0] += sh[2][w];
            sh[3][0] += sh[3][w];
            sh[4][0] = fmin(sh[4][0], sh[4][w]);
            sh[5][0] = fmax(sh[5][0], sh[5][w]);
        }
        double* o = part + ((size_t)j * gridDim.x + blockIdx.x) * 6;
        for (int q = 0; q < 6; ++q) o[q] = sh[q][0];
    }
}

// K7: block partials -> per-column results of all ranks in ONE exchange (min part | sum part), then the packed buffer
//   small = [status, 0 x 7 | per column: ds, mad, min, max, sum, n_acc, median(x), accepted statistic constant?]
__global__ void __launch_bounds__(256) k_cols_final(PeerDev PD, int world, const double* __restrict__ part, int gx, int P,
                                                    const double* __restrict__ sel, const double* __restrict__ mad,
                                                    const double* __restrict__ med, const int* __restrict__ status,
                                                    double* __restrict__ small) {
    extern __shared__ double vec[];   // [P min | P -max | P xmin | P -xmax | 8 -status bits || P sum | P count]
    const int tid = threadIdx.x;
    const int n_min = 4 * P + 8, count = n_min + 2 * P;
    int st = *status;
    if (world > 1 && ld_volatile_u64(PD.err) != 0ull) st |= ABCDLS_S_PEER_TIMEOUT;
    if (tid < P) {
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        double mn = inf, mx = -inf, sum = 0.0, num = 0.0, xmn = inf, xmx = -inf;
        for (int b = 0; b < gx; ++b) {           // block order: deterministic
            const double* o = part + ((size_t)tid * gx + b) * 6;
            mn = fmin(mn, o[0]);
            mx = fmax(mx, o[1]);
            sum += o[2];
            num += o[3];
            xmn = fmin(xmn, o[4]);
            xmx = fmax(xmx, o[5]);
        }
        vec[tid] = mn;
        vec[P + tid] = -mx;
        vec[2 * P + tid] = xmn;
        vec[3 * P + tid] = -xmx;
        vec[n_min + tid] = sum;
        vec[n_min + P + tid] = num;
    }
    if (tid < 8) vec[4 * P + tid] = -(double)((st >> tid) & 1);
    __syncthreads();
    if (world > 1) {
        peer_block_allreduce(PD, vec, count, n_min);
        if (tid == 0 && ld_volatile_u64(PD.err) != 0ull) vec[4 * P + 6] = -1.0;
        __syncthreads();
    }
    if (tid < P) {
        double* o = small + 8 + (size_t)tid * 8;
        o[0] = sel[2 * tid];
        o[1] = mad[tid];
        o[2] = vec[tid];
        o[3] = -vec[P + tid];
        o[4] = vec[n_min + tid];
        o[5] = vec[n_min + P + tid];
        o[6] = med[tid];
        o[7] = (vec[2 * P + tid] == -vec[3 * P + tid]) ? 1.0 : 0.0;
    }
    if (tid == 0) {
        int so = 0;
        for (int b = 0; b < 8; ++b)
            if (vec[4 * P + b] != 0.0) so |= 1 << b;
        small[0] = (double)so;
        for (int q = 1; q < 8; ++q) small[q] = 0.0;
    }
}

}  // namespace

extern "C" {

int abcdls_columns_supported(const double* ss, int64_t n, int P, int64_t ld) {
    return ss && n >= (1 << 18) && n < (1ll << 31) && P >= 1 && P <= kMaxD && ld >= n;
}

size_t abcdls_columns_workspace_bytes(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint) {
    if (!h) return 0;
    return carve_cols(nullptr, P, n, cols_capT(n, nacc_local_hint), h->sm_count).total;
}

// Stage 1: column samples.  est_out: doubles a multi-rank caller all-reduces (sum) before stage 2.
int abcdls_columns_sample(abcdls_handle_t h, const double* ss, int64_t n, int P, int64_t ld, const double* target,
                          int64_t nacc, int64_t n_global, int64_t nacc_local_hint, int world, void* ws, size_t ws_bytes,
                          double** est_out, size_t* est_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && target && ws && abcdls_columns_supported(ss, n, P, ld) && world >= 1 && nacc >= 1 &&
                            n_global >= nacc, "columns_sample");
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    if (ws_bytes < w.total) {
        h->err = "columns workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const double delta = col_delta(n, world);
    const int Gw = sample_ctas_w(n, world);
    // (p + 5.5 sigma) quantile of |x - t|, sigma of a sample quantile over the Gw * 4096 * world draws of all ranks
    const double p = (double)nacc / (double)n_global;
    const double stot = (double)Gw * kSample * world;
    double pT = p + kZ * sqrt(p * (1.0 - p) / stot) + 2.0 / kSample;
    if (pT > 1.0) pT = 1.0;
    const size_t smem = (size_t)2 * kSample * sizeof(double);
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_sample_cols_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.m.cnt, 0, w.m.zero_bytes, st));
    ABCDLS_CUDA(h, cudaMemsetAsync(w.cntT, 0, (size_t)((char*)w.valT - (char*)w.cntT), st));   // cntT | red | cnt2
    double* estT = w.m.est + (size_t)Gw * P * 4;
    k_sample_cols_t<<<dim3(Gw, P), 512, smem, st>>>(ss, n, ld, 0x5EEDu, delta, target, pT, w.m.est, estT);
    ABCDLS_LAUNCH_CHECK(h);
    if (est_out) *est_out = w.m.est;
    if (est_count) *est_count = (size_t)Gw * P * 5;
    return ABCDLS_OK;
}

// Stage 2: brackets + THE pass.  counts_out: int64 block (counts | gcnt | histM) to all-reduce (sum); tcount_out: the
// P member counts (int64) to all-reduce (sum) as well.
int abcdls_columns_pass(abcdls_handle_t h, const double* ss, int64_t n, int P, int64_t ld, const double* target,
                        int64_t nacc_local_hint, int world, void* ws, size_t ws_bytes, int64_t** counts_out,
                        size_t* counts_count, int64_t** tcount_out, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && target && ws && abcdls_columns_supported(ss, n, P, ld) && world >= 1, "columns_pass");
    const long long capM = cap_for(n, 0.03), capU = cap_for(n, 0.08);
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    if (ws_bytes < w.total) {
        h->err = "columns workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int Gw = sample_ctas_w(n, world);
    k_make_brackets<<<P, 32, 0, st>>>(w.m.est, Gw, P, 1.0 / world, w.m.brk);
    k_make_tband<<<P, 32, 0, st>>>(w.m.est + (size_t)Gw * P * 4, Gw, P, 1.0 / world, w.RT);
    ABCDLS_LAUNCH_CHECK(h);
    long long gx = ((n + kATile - 1) / kATile);
    long long mx = (long long)h->sm_count * 3 / P;    // three resident CTAs per SM (75 KB of shared memory each)
    if (mx < 1) mx = 1;
    if (gx > mx) gx = mx;
    const size_t smem = (size_t)kNB * 4 + (size_t)(kAThreads / 32) * ((kStageM + kStageU + kCStageT) * 8 + kCStageT * 4);
    ABCDLS_ONCE(h, cudaFuncSetAttribute(k_columns_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_columns_pass<<<dim3((unsigned)gx, P), kAThreads, smem, st>>>(ss, n, ld, w.m.brk, target, w.RT, w.m.counts, w.m.histM,
                                                                  w.m.candM, capM, w.m.candU, capU, w.m.cnt, w.valT,
                                                                  w.rowT, w.capT, w.cntT);
    k_copy_u64<<<1, 128, 0, st>>>(w.m.gcnt, w.m.cnt, 2 * P);
    k_copy_u64<<<1, 128, 0, st>>>(w.red, w.cntT, P);
    ABCDLS_LAUNCH_CHECK(h);
    if (counts_out) *counts_out = reinterpret_cast<int64_t*>(w.m.counts);
    if (counts_count) *counts_count = w.m.redA_count;
    if (tcount_out) *tcount_out = reinterpret_cast<int64_t*>(w.red);
    return ABCDLS_OK;
}

// Stage 3: medians / MADs of all columns: the phases and exchanges of abcdls_mad_fast_refine, on this workspace.
int abcdls_columns_refine(abcdls_handle_t h, int64_t n, int64_t n_global, int P, int64_t nacc_local_hint, int phase,
                          int world, const double* gath_small, const int32_t* gath_cnt, void* ws, size_t ws_bytes,
                          double* med, double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1,
                          size_t* x1_count, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && med && mad && status_dev && P >= 1 && P <= kMaxD && phase >= 0 && phase <= 3,
                     "columns_refine");
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    return refine_impl(h, w.m, cap_for(n, 0.03), cap_for(n, 0.08), false, n_global, P, phase, world, gath_small,
                       gath_cnt, med, mad, status_dev, x0, x0_count, x1, x1_count, (cudaStream_t)stream);
}

// Stage 4: R's exact distance of every member.  dist_out / cnt_out: the (P, capT) table and per-column counts the
// caller hands to abcdls_select_* (rank nacc - 1 over all ranks -> sel[2 j] = ds_j).
int abcdls_columns_exact(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint, const double* mad,
                         const double* target, void* ws, size_t ws_bytes, double** dist_out, int64_t* cap_out,
                         int64_t* cnt_out, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && mad && target && ws && cnt_out && status_dev && P >= 1 && P <= kMaxD, "columns_exact");
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    unsigned gx = (unsigned)std::min<long long>((w.capT + 255) / 256, (long long)std::max(1, h->sm_count * 4 / P));
    k_cols_exact<<<dim3(gx, P), 256, 0, (cudaStream_t)stream>>>(w.valT, w.capT, w.cntT, mad, target, w.distT,
                                                                (long long*)cnt_out, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    if (dist_out) *dist_out = w.distT;
    if (cap_out) *cap_out = w.capT;
    return ABCDLS_OK;
}

// Stage 5: counts around ds + the completeness proof.  cnt2_out: [P][2] int64 {n_lt, n_eq} of this rank (all-gather).
int abcdls_columns_count(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint, int64_t nacc, const double* sel,
                         const double* mad, const double* target, const int64_t* cnt, void* ws, size_t ws_bytes,
                         int64_t** cnt2_out, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && sel && mad && target && cnt && ws && status_dev && P >= 1 && P <= kMaxD, "columns_count");
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    unsigned gx = (unsigned)std::min<long long>((w.capT + 255) / 256, (long long)std::max(1, h->sm_count * 4 / P));
    k_cols_count<<<dim3(gx, P), 256, 0, (cudaStream_t)stream>>>(w.distT, w.capT, (const long long*)cnt, sel, mad, target,
                                                                w.RT, w.red, nacc, w.cnt2, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    if (cnt2_out) *cnt2_out = reinterpret_cast<int64_t*>(w.cnt2);
    return ABCDLS_OK;
}

// Stage 6a: row keys of the cap rule.  keys_out: (P, cap) table for abcdls_select_* (counts = cnt, NO cross-rank
// reduction): rank quota_j - 1 gives thr_j, the last local row of column j this rank accepts (quota_j = how many of its
// n_lt + n_eq members with dist <= ds this rank may keep once the lower ranks took theirs; 0 -> pass thr_j = -1).
int abcdls_columns_rowkeys(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint, const double* sel,
                           const int64_t* cnt, void* ws, size_t ws_bytes, double** keys_out, void* stream) {
    ABCDLS_CHECK_ARG(h, h && sel && cnt && ws && P >= 1 && P <= kMaxD, "columns_rowkeys");
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    unsigned gx = (unsigned)std::min<long long>((w.capT + 255) / 256, (long long)std::max(1, h->sm_count * 4 / P));
    k_cols_rowkeys<<<dim3(gx, P), 256, 0, (cudaStream_t)stream>>>(w.distT, w.rowT, w.capT, (const long long*)cnt, sel,
                                                                  w.keys);
    ABCDLS_LAUNCH_CHECK(h);
    if (keys_out) *keys_out = w.keys;
    return ABCDLS_OK;
}

// Stage 6b: accepted parameters + the packed result.  thr[2 j] (device doubles, layout of abcdls_select_end): members
// with dist <= ds_j and local row <= thr[2 j] are accepted.  par: element (row i, column j) at
// par[j * par_col_stride + i * par_row_stride].
// acc_* (may be NULL): the accepted (local row, parameter, distance) of every column, unordered, acc_cap per column.
// small_out[8 + 8 P]: {status, 0 x 7 | per column: ds, mad, min, max, sum, n_acc, median of the statistic, accepted
// statistic constant?}.
int abcdls_columns_accept(abcdls_handle_t h, abcdls_peer_t peer, int64_t n, int P, int64_t nacc_local_hint,
                          const double* sel, const double* med, const double* mad, const int64_t* cnt,
                          const double* thr, const double* par, int64_t par_col_stride, int64_t par_row_stride,
                          int64_t acc_cap, int32_t* acc_row, double* acc_val, double* acc_dist, int64_t* acc_cnt,
                          void* ws, size_t ws_bytes, double* small_out, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && sel && med && mad && cnt && thr && par && ws && small_out && status_dev && P >= 1 &&
                            P <= kMaxD, "columns_accept");
    ABCDLS_CHECK_ARG(h, !acc_row || (acc_val && acc_dist && acc_cnt && acc_cap >= 1), "columns_accept: lists");
    ColsWs w = carve_cols(ws, P, n, cols_capT(n, nacc_local_hint), h->sm_count);
    cudaStream_t st = (cudaStream_t)stream;
    if (acc_cnt) ABCDLS_CUDA(h, cudaMemsetAsync(acc_cnt, 0, (size_t)P * 8, st));
    k_cols_accept<<<dim3(w.gx_acc, P), 256, 0, st>>>(w.distT, w.rowT, w.valT, w.capT, (const long long*)cnt, sel, thr, par,
                                                     par_col_stride, par_row_stride, w.part, acc_cap, acc_row, acc_val,
                                                     acc_dist, (unsigned long long*)acc_cnt);
    PeerDev PD = {};
    if (peer) PD = peer->dev;
    k_cols_final<<<1, 256, (size_t)(6 * P + 8) * sizeof(double), st>>>(PD, peer ? peer->dev.world : 1, w.part, w.gx_acc, P,
                                                                      sel, mad, med, status_dev, small_out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
