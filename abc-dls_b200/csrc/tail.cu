// tail.cu — the tail of abc() on the accepted rows as THREE launches (two for rejection), each ending in its own
// reduction, cross-rank exchange and small solve:
//
//   T1 k_tail_gather_fit    accepted rows (candidate block / table, row- or column-major parameters) -> compact
//                           column-major blocks xs | y | ss | w, written once; from the same shared-memory tile the
//                           weighted normal equations X'WX | X'WY, the unweighted column sums and the region min / max.
//                           Last block: partials -> [NVLink mailbox all-reduce] -> Cholesky -> beta.
//   T2 k_tail_resid_fit     (hcorr only) residuals from xs, y (not materialised) -> R's second fit X'W log(res^2).
//                           Last block: -> [all-reduce] -> gamma (stored Cholesky factor).
//   T3 k_tail_adjust_final  adj = pred + the_m + res * exp(-gamma.x / 2), residuals, summary statistics of the posterior
//                           sample, sum w res^2 (sigma2).  Last block: ONE exchange (min part | sum part) -> the packed
//                           `small` result buffer.
// The normal equations run on the fp64 tensor-core path (mma.sync m8n8k4: 256 FMAs per issued instruction - at 561
// FMAs per accepted row the scalar version was issue bound); everything else is plain fp64.
// Measured dead ends (B200, 1e6 accepted rows, d = P = 16; profiles/r02_notes.md): T1 stays at ~0.30 ms whether it
// runs as 2 x 256-thread CTAs per SM, as 4 x 128, warp specialised with a double-buffered tile (8 loader + 4 / 8
// accumulator warps), with DMMA or with register-blocked DFMA - it is bound by the 2e6 RANDOM 128-byte record reads
// (one DRAM row activation each: the accepted rows are 1 % of the table, and the candidate block is in pass order).
// Round-2 elimination experiment (parts of T1 switched off by a debug mask, 1e6 rows, T1 alone): 303 us in all =
// ~117 us skeleton (row numbers, Markstein scaling, tile stores, barriers, last-block reduction + Cholesky) + ~64 us
// record reads (2e6 random lines: ~31 G lines/s) + ~55 us stores (344 MB at the DRAM write rate) + ~62 us DMMA + ~12 us
// region scan: the parts ADD UP (two CTAs per SM in phase) instead of overlapping.  Prefetching the records one tile
// ahead with cp.async (thread per record, or 8 lanes per record) and splitting both records between the two half
// blocks changed the sum by less than 7 %; those variants are not kept.
//
// Reference semantics: CRAN abc 2.2.1 abc(method = "loclinear" | "rejection") as reached from
// src/Classes/ABC.py:1950-1953 / 2806-2809; same arithmetic as loclinear.cu (design centred on
// the scaled target, fp64 normal equations, fixed reduction order) with one change: R's unweighted recentring
// the_m = colMeans(residuals) is obtained algebraically from column sums of T1,
//     the_m = mean(y) - beta0 - sum_j beta_j mean(xs_j),
// so that the residuals need no pass of their own (error ~ eps |mean y|, far inside the 1e-9 bar).
// This replaces 16 launches + 3 stand-alone exchanges of loclinear.cu (which stays as the path for d or P > 32 and for
// the NCCL fallback).  DESIGN.md §3.4.
#include "peer.cuh"

#include <algorithm>

namespace {

constexpr int kTlRows = 128;              // accepted rows per shared-memory tile
constexpr int kTlPitch = 132;             // [column][row] tiles: DMMA fragment loads (8 columns x 4 rows) hit 16 banks
constexpr int kTlThreads = 256;
constexpr int kTlWarps = kTlThreads / 32;
constexpr int kMA = 5, kMB = 3;           // 8 x 8 output tiles per warp of the normal equations: up to 5 x 3
constexpr int kMT = kMA * kMB;
constexpr int kTlGroup = 16;              // CTAs per first-level reduction group
constexpr int kTlMaxD = 32, kTlMaxP = 32;

struct TailWs {
    double* part;     // [G][E]        CTA partials (the three kernels run back to back on one stream and reuse it)
    double* gpart;    // [groups][E]
    double* vecA;     // gram M*M | rhs M*P | usum W  (summed over ranks): usum = {n, sum xs_j, sum y_p}, vecA[0] = sum w
    double* reg;      // region min (d) | region max (d) of the accepted statistics, this rank
    double* L;        // Cholesky factor of the gram (M*M, lower)
    double* beta;     // M*P   (row 0 = intercept = prediction at the target)
    double* gamma;    // M*P
    unsigned* tickets;   // 3 x (maxgroups + 1), zero between launches
    int maxgroups;
};

struct TailArgs {
    // where the accepted rows come from
    const double* cand; long long cand_pitch; const long long* slot; const long long* perm;   // one-pass path
    const double* ss; long long ld_ss;                                                        // otherwise: the table
    const double* par_cm; long long ld_par;      // column-major parameter table, or
    const double* par_rm; long long par_pitch;   // row-major (device or pinned host)
    const double* dacc; const long long* idx; const long long* n_acc; long long row_offset; long long cap;
    const double* mad; const double* target; const double* ds;
    // compact blocks of the accepted rows (column-major, leading dimension cap)
    double* xs; double* y; double* ss_out; double* w; double* adj; double* res;
    int d, P, lin, hcorr, world;
    int* status;
    double* small_out;
    TailWs ws;
};

__device__ __forceinline__ double ldcg_f64(const double* p) { return __ldcg(p); }

// src[0], src[stride], ... src[(cnt - 1) * stride] combined IN ORDER (kd: 0 sum, 1 min, 2 max), eight loads in flight at
// a time: one L2 round trip per eight partials instead of one per partial.
__device__ __forceinline__ double ordered_combine(const double* src, size_t stride, int cnt, int kd) {
    double acc = kd == 0 ? 0.0 : (kd == 1 ? __longlong_as_double(0x7ff0000000000000ll)
                                          : -__longlong_as_double(0x7ff0000000000000ll));
    for (int c0 = 0; c0 < cnt; c0 += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = c0 + k < cnt ? ldcg_f64(src + (size_t)(c0 + k) * stride) : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c0 + k < cnt) acc = (c0 + k == 0) ? v[k] : (kd == 0 ? acc + v[k] : (kd == 1 ? fmin(acc, v[k]) : fmax(acc, v[k])));
    }
    return acc;
}

// Every CTA has written E doubles to part[blockIdx.x * E ...].  Two-level, fixed-order combine: the last CTA of each
// group of kTlGroup adds the group's partials in CTA order, the last group to finish adds the group sums in group
// order (deterministic for a given grid).  Returns true in exactly one CTA, with the result in out[0..E).
// kind(e): 0 sum, 1 min, 2 max.
template <class Kind>
__device__ bool hier_reduce(const double* part, int E, double* gpart, unsigned* tickets, double* out, Kind kind) {
    __shared__ int role;
    const int G = gridDim.x, ngroups = (G + kTlGroup - 1) / kTlGroup, g = blockIdx.x / kTlGroup;
    const int gsize = min(kTlGroup, G - g * kTlGroup);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) role = (atomicAdd(&tickets[g], 1u) == (unsigned)gsize - 1u) ? 1 : 0;
    __syncthreads();
    if (role == 0) return false;
    __threadfence();
    for (int e = threadIdx.x; e < E; e += blockDim.x)
        gpart[(size_t)g * E + e] = ordered_combine(part + (size_t)g * kTlGroup * E + e, E, gsize, kind(e));
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        tickets[g] = 0u;
        role = (atomicAdd(&tickets[ngroups], 1u) == (unsigned)ngroups - 1u) ? 2 : 0;
    }
    __syncthreads();
    if (role != 2) return false;
    __threadfence();
    for (int e = threadIdx.x; e < E; e += blockDim.x) out[e] = ordered_combine(gpart + e, E, ngroups, kind(e));
    if (threadIdx.x == 0) tickets[ngroups] = 0u;
    __syncthreads();
    return true;
}

// In-place Cholesky of the symmetric M x M matrix in L (row-major; the lower triangle is used), right-looking so that
// every step is one parallel sweep.  Same pivot rule as loclinear.cu:k_solve (lsfit's "deemed singular").  All threads
// of the block call it; *bad (shared) is set on a rejected pivot.  diag0: M doubles of scratch.
__device__ void chol_factor(double* L, int M, double* diag0, int* bad) {
    for (int a = threadIdx.x; a < M; a += blockDim.x) diag0[a] = L[a * M + a];
    if (threadIdx.x == 0) *bad = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double dmax = 0.0;
        for (int a = 0; a < M; ++a) dmax = fmax(dmax, fabs(diag0[a]));
        if (!(dmax > 0.0)) *bad = 1;
    }
    for (int c = 0; c < M; ++c) {
        __syncthreads();
        // every thread takes the pivot itself (no one-thread section, two barriers per column)
        double s = L[c * M + c];
        if (!(s > 1e-13 * diag0[c]) || !(s > 0.0)) {
            if (threadIdx.x == 0) *bad = 1;
            s = 1.0;
        }
        const double lcc = sqrt(s);
        double below = 0.0;
        const int r1 = c + 1 + threadIdx.x;
        if (r1 < M) below = L[r1 * M + c] / lcc;                      // M <= 33 <= blockDim.x: one row per thread
        __syncthreads();                                              // everyone has read the pivot and its column
        if (threadIdx.x == 0) L[c * M + c] = lcc;
        if (r1 < M) L[r1 * M + c] = below;
        __syncthreads();
        const int nn = M - c - 1;
        for (int t = threadIdx.x; t < nn * nn; t += blockDim.x) {
            const int r = c + 1 + t / nn, k = c + 1 + t % nn;
            if (k <= r) L[r * M + k] = fma(-L[r * M + c], L[k * M + c], L[r * M + k]);
        }
    }
    __syncthreads();
}

// L L' X = rhs for P right-hand sides (rhs, out: M x P row-major); Z: M x P doubles of shared scratch.
__device__ void chol_solve(const double* L, int M, const double* rhs, int P, double* Z, double* out, int bad) {
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        for (int a = 0; a < M; ++a) {
            double t = rhs[a * P + p];
            for (int k = 0; k < a; ++k) t = fma(-L[a * M + k], Z[k * P + p], t);
            Z[a * P + p] = t / L[a * M + a];
        }
        for (int a = M - 1; a >= 0; --a) {
            double t = Z[a * P + p];
            for (int k = a + 1; k < M; ++k) t = fma(-L[k * M + a], Z[k * P + p], t);
            Z[a * P + p] = t / L[a * M + a];
        }
        for (int a = 0; a < M; ++a) out[a * P + p] = bad ? __longlong_as_double(0x7ff8000000000000ll) : Z[a * P + p];
    }
    __syncthreads();
}

// ---- C[a][b] += sum_r (w_r T_r[a]) T_r[b] over one staged tile, on the fp64 tensor-core path ---------------------------
// mma.sync.m8n8k4.f64: A = 8 output rows a x 4 table rows r (lane l: a = l / 4, r = l % 4), B = 4 r x 8 output columns b
// (lane l: r = l % 4, b = l / 4), C = 8 x 8 (lane l: a = l / 4, b = 2 (l % 4) + {0, 1}).  256 FMAs per issued
// instruction instead of 32: the normal equations stop being issue bound (561 FMAs per accepted row at d = P = 16).
// A warp owns up to kMA x kMB output tiles in registers; the warps split into nsl groups by tile column, the warps of a
// group share the k-steps (4 table rows each) of the tile.  Output row a == M (T1 only) is the UNWEIGHTED ones row: its
// tile row holds the plain column sums {n, sum xs_j, sum y_p}.  The accumulation order is fixed: deterministic.
template <int NA>
struct MmaMap {
    int nAt, nBt, nsl, b0, rowsA;
    int aoff[NA], boff[kMB];
    bool awm[NA];
    int ks0, ksstep;
};

// NA x kMB register tiles per warp; rows / columns beyond the real shape point at the all-zero tile column, so the inner
// loop has no guards (a padded tile is a wasted DMMA, never a wrong one).
template <int NA>
__device__ __forceinline__ MmaMap<NA> make_mma_map(int M, int W, int b0, bool ones_row) {
    MmaMap<NA> m;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    m.b0 = b0;
    m.rowsA = M + (ones_row ? 1 : 0);
    m.nAt = (m.rowsA + 7) / 8;
    m.nBt = (W - b0 + 7) / 8;
    int nsl = 1;
    while ((m.nBt + nsl - 1) / nsl > kMB) nsl <<= 1;
    m.nsl = nsl;
    const int g = wid % nsl;
#pragma unroll
    for (int ta = 0; ta < NA; ++ta) {
        const int a = 8 * ta + (lane >> 2);
        m.aoff[ta] = a < M ? a * kTlPitch : ((ones_row && a == M) ? 0 : (W + 1) * kTlPitch);
        m.awm[ta] = a < M;
    }
#pragma unroll
    for (int tb = 0; tb < kMB; ++tb) {
        const int b = b0 + 8 * (g + nsl * tb) + (lane >> 2);
        m.boff[tb] = (g + nsl * tb < m.nBt && b < W) ? b * kTlPitch : (W + 1) * kTlPitch;
    }
    m.ks0 = wid / nsl;
    m.ksstep = kTlWarps / nsl;
    return m;
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int NA>
__device__ __forceinline__ void mma_accumulate(const double* __restrict__ T, int W, const MmaMap<NA>& m,
                                               double (&C)[NA * kMB][2]) {
    const int lane = threadIdx.x & 31;
    const double* __restrict__ wt = T + (size_t)W * kTlPitch + (lane & 3);
    const double* __restrict__ Tl = T + (lane & 3);
#pragma unroll 2
    for (int ks = m.ks0; ks < kTlRows / 4; ks += m.ksstep) {
        const int r = 4 * ks;
        const double wv = wt[r];
        double af[NA], bf[kMB];
#pragma unroll
        for (int ta = 0; ta < NA; ++ta) {
            const double v = Tl[m.aoff[ta] + r];
            af[ta] = m.awm[ta] ? v * wv : v;
        }
#pragma unroll
        for (int tb = 0; tb < kMB; ++tb) bf[tb] = Tl[m.boff[tb] + r];
#pragma unroll
        for (int ta = 0; ta < NA; ++ta)
#pragma unroll
            for (int tb = 0; tb < kMB; ++tb) dmma(C[ta * kMB + tb], af[ta], bf[tb]);
    }
}

// warps -> one CTA partial, warps added in order.  red: kTlWarps * NA * kMB * 64 doubles of shared memory.
// out layout: C[a][b] at (b < M ? a * M + b : offY + a * P + (b - M)); the ones row (a == M) at offU + b.
template <int NA>
__device__ void mma_combine(double* red, const MmaMap<NA>& m, const double (&C)[NA * kMB][2], int M, int W, int P,
                            int offY, int offU, double* out) {
    constexpr int NT = NA * kMB;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        double* dst = red + ((size_t)(wid * NT + t) * 64 + lane * 2);
        dst[0] = C[t][0];
        dst[1] = C[t][1];
    }
    __syncthreads();
    const int Wb = W - m.b0;
    for (int t = threadIdx.x; t < m.rowsA * Wb; t += blockDim.x) {
        const int a = t / Wb, bb = t - a * Wb, b = m.b0 + bb;
        const int ta = a >> 3, tb = bb >> 3, g = tb % m.nsl, tbl = tb / m.nsl;
        const int pos = (a & 7) * 8 + (bb & 7);
        double sum = 0.0;
        for (int wq = g; wq < kTlWarps; wq += m.nsl) sum += red[(size_t)(wq * NT + ta * kMB + tbl) * 64 + pos];
        const int e = a < M ? ((b < M) ? a * M + b : offY + a * P + (b - M)) : offU + b;
        out[e] = sum;
    }
    __syncthreads();
}

__device__ __forceinline__ size_t region0_doubles(int d, int P) {
    const size_t tile = (size_t)(d + 1 + P + 2 + d) * kTlPitch;
    const size_t red = (size_t)kTlWarps * kMT * 64;
    return tile > red ? tile : red;
}

// x / div correctly rounded (== R's x / mad): Markstein with the reciprocal while everything stays in the normal range
// (csrc/common.cuh:div_const), IEEE division otherwise.  rcp == 0 marks a divisor outside [1e-90, 1e90].
__device__ __forceinline__ double scaled_by(double x, double div, double rcp) {
    const double ax = fabs(x);
    if ((ax > 1e-200 && ax < 1e200 && rcp != 0.0) || x == 0.0) return div_const(x, div, rcp);
    return x / div;
}

// =====================================================================================================================
// T1.  Two threads per accepted row: thread (rr, 0) reads the row's statistics (one 8 d-byte record of the candidate
// block, or d scattered table cells), scales them and writes xs / ss straight to the compact blocks; thread (rr, 1) does
// the same for the parameters (one 8 P-byte record of a row-major table).  Both leave a copy in the [column][row] tile
// for the tensor-core accumulation.  The row numbers of the NEXT tile are fetched while this one accumulates.
template <int NA>
__global__ void __launch_bounds__(kTlThreads, 2) k_tail_gather_fit(TailArgs A, PeerDev PD) {
    pdl_prologue();
    extern __shared__ __align__(16) double sm[];
    const int d = A.d, P = A.P, M = d + 1, W = M + P, lin = A.lin;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* T = sm;                                        // [W + 2][pitch]: ones | xs | y | w | zeros
    double* R = T + (size_t)(W + 2) * kTlPitch;            // [d][pitch]: raw statistics
    double* red = sm;                                      // aliases T / R after the tile loop
    double* sdiv = sm + region0_doubles(d, P);             // [d]
    double* srcp = sdiv + d;                               // [d]
    double* stgt = srcp + d;                               // [d]
    __shared__ int bad;

    for (int j = tid; j < d; j += kTlThreads) {
        const double m = A.mad[j];
        const double dv = (m == 0.0) ? 1.0 : m;
        const double av = fabs(dv);
        sdiv[j] = dv;
        srcp[j] = (av > 1e-90 && av < 1e90) ? 1.0 / dv : 0.0;
        stgt[j] = (m == 0.0) ? A.target[j] : A.target[j] / dv;
    }
    for (int t = tid; t < kTlPitch; t += kTlThreads) T[(size_t)(W + 1) * kTlPitch + t] = 0.0;
    long long n_acc = *A.n_acc;
    if (n_acc > A.cap) n_acc = A.cap;
    const double ds = *A.ds;
    const long long ntile = (n_acc + kTlRows - 1) / kTlRows;
    const long long per = (ntile + gridDim.x - 1) / gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per, t_end = min(ntile, t_begin + per);

    const MmaMap<NA> mm = make_mma_map<NA>(M, W, 0, true);
    double C[NA * kMB][2];
#pragma unroll
    for (int t = 0; t < NA * kMB; ++t) C[t][0] = C[t][1] = 0.0;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double rmin[4] = {inf, inf, inf, inf}, rmax[4] = {-inf, -inf, -inf, -inf};   // columns wid, wid + 8, ...
    const int rr = tid & (kTlRows - 1), h = tid >> 7;
    const bool cand_vec = A.cand && !(d & 1) && !(A.cand_pitch & 1) && !(reinterpret_cast<uintptr_t>(A.cand) & 15);
    const bool par_vec = A.par_rm && !(P & 1) && !(A.par_pitch & 1) && !(reinterpret_cast<uintptr_t>(A.par_rm) & 15);

    // row r -> (index of its statistics record / table row, distance) for h == 0, its parameter row for h == 1
    auto fetch_row = [&](long long r, long long& src, double& q) {
        src = 0;
        q = 0.0;
        if (r >= n_acc) return;
        if (h == 0) {
            if (A.cand) {
                src = A.slot[r];
                if (A.perm) src = A.perm[src];
            } else {
                src = A.idx[r] - A.row_offset;
            }
            q = A.dacc[r];
        } else {
            src = A.idx[r] - A.row_offset;
        }
    };
    long long src_cur;
    double q_cur;
    fetch_row(t_begin * kTlRows + rr, src_cur, q_cur);
    __syncthreads();

    for (long long tile = t_begin; tile < t_end; ++tile) {
        const long long r0 = tile * kTlRows;
        const int nr = (int)min((long long)kTlRows, n_acc - r0);
        const bool valid = rr < nr;
        const size_t r = (size_t)(r0 + rr);
        if (h == 0) {
            // ---- statistics: raw -> ss_out, R; scaled and centred on the scaled target -> xs, T[1 + j] ----
            // the record is fetched in chunks of 8 values: the loads of a chunk are in flight together
            if (cand_vec) {
                const double2* rec = reinterpret_cast<const double2*>(A.cand + (size_t)src_cur * A.cand_pitch);
                for (int j0 = 0; j0 < d; j0 += 8) {
                    double2 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[k] = make_double2(0.0, 0.0);
                        if (valid && j0 + 2 * k < d) v[k] = rec[(j0 >> 1) + k];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = j0 + 2 * k;
                        if (j < d) {
                            const double x0 = __dsub_rn(scaled_by(v[k].x, sdiv[j], srcp[j]), stgt[j]);
                            const double x1 = __dsub_rn(scaled_by(v[k].y, sdiv[j + 1], srcp[j + 1]), stgt[j + 1]);
                            R[(size_t)j * kTlPitch + rr] = v[k].x;
                            R[(size_t)(j + 1) * kTlPitch + rr] = v[k].y;
                            T[(size_t)(1 + j) * kTlPitch + rr] = valid ? x0 : 0.0;
                            T[(size_t)(2 + j) * kTlPitch + rr] = valid ? x1 : 0.0;
                            if (valid) {
                                A.ss_out[(size_t)j * A.cap + r] = v[k].x;
                                A.ss_out[(size_t)(j + 1) * A.cap + r] = v[k].y;
                                if (lin) {
                                    A.xs[(size_t)j * A.cap + r] = x0;
                                    A.xs[(size_t)(j + 1) * A.cap + r] = x1;
                                }
                            }
                        }
                    }
                }
            } else {
                for (int j0 = 0; j0 < d; j0 += 8) {
                    double v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int j = j0 + k;
                        v[k] = 0.0;
                        if (valid && j < d)
                            v[k] = A.cand ? A.cand[(size_t)src_cur * A.cand_pitch + j] : A.ss[(size_t)j * A.ld_ss + src_cur];
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int j = j0 + k;
                        if (j < d) {
                            const double x0 = __dsub_rn(scaled_by(v[k], sdiv[j], srcp[j]), stgt[j]);
                            R[(size_t)j * kTlPitch + rr] = v[k];
                            T[(size_t)(1 + j) * kTlPitch + rr] = valid ? x0 : 0.0;
                            if (valid) {
                                A.ss_out[(size_t)j * A.cap + r] = v[k];
                                if (lin) A.xs[(size_t)j * A.cap + r] = x0;
                            }
                        }
                    }
                }
            }
            double wv = valid ? 1.0 : 0.0;
            if (valid && lin) {
                const double q = q_cur / ds;
                wv = __dsub_rn(1.0, __dmul_rn(q, q));       // Epanechnikov, R: 1 - (dist / ds)^2
                A.w[r] = wv;
            }
            T[(size_t)W * kTlPitch + rr] = wv;
            T[rr] = valid ? 1.0 : 0.0;                       // the ones column (also counts the rows)
        } else {
            // ---- parameters -> y, T[M + p] ----
            if (par_vec) {
                const double2* rec = reinterpret_cast<const double2*>(A.par_rm + (size_t)src_cur * A.par_pitch);
                for (int p0 = 0; p0 < P; p0 += 8) {
                    double2 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[k] = make_double2(0.0, 0.0);
                        if (valid && p0 + 2 * k < P) v[k] = rec[(p0 >> 1) + k];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int p = p0 + 2 * k;
                        if (p < P) {
                            T[(size_t)(M + p) * kTlPitch + rr] = v[k].x;
                            T[(size_t)(M + p + 1) * kTlPitch + rr] = v[k].y;
                            if (valid) {
                                A.y[(size_t)p * A.cap + r] = v[k].x;
                                A.y[(size_t)(p + 1) * A.cap + r] = v[k].y;
                            }
                        }
                    }
                }
            } else {
                for (int p0 = 0; p0 < P; p0 += 8) {
                    double v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int p = p0 + k;
                        v[k] = 0.0;
                        if (valid && p < P)
                            v[k] = A.par_rm ? A.par_rm[(size_t)src_cur * A.par_pitch + p]
                                            : A.par_cm[(size_t)p * A.ld_par + src_cur];
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int p = p0 + k;
                        if (p < P) {
                            T[(size_t)(M + p) * kTlPitch + rr] = v[k];
                            if (valid) A.y[(size_t)p * A.cap + r] = v[k];
                        }
                    }
                }
            }
        }
        long long src_next;
        double q_next;
        fetch_row(tile + 1 < t_end ? r0 + kTlRows + rr : n_acc, src_next, q_next);   // in flight during the accumulation
        __syncthreads();
        // ---- region check: min / max of the raw statistics ----
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = wid + kTlWarps * q;
            if (j < d)
                for (int i = lane; i < nr; i += 32) {
                    const double x = R[(size_t)j * kTlPitch + i];
                    rmin[q] = fmin(rmin[q], x);
                    rmax[q] = fmax(rmax[q], x);
                }
        }
        if (lin) mma_accumulate<NA>(T, W, mm, C);
        src_cur = src_next;
        q_cur = q_next;
        __syncthreads();
    }

    // ---- CTA partial: [gram | X'WY | column sums] (lin) | region min | region max ----
    const int offY = M * M, offU = M * M + M * P;
    const int off_reg = lin ? offU + W : 0;
    const int E = off_reg + 2 * d;
    double* mypart = A.ws.part + (size_t)blockIdx.x * E;
    if (lin) mma_combine<NA>(red, mm, C, M, W, P, offY, offU, mypart);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = wid + kTlWarps * q;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rmin[q] = fmin(rmin[q], __shfl_xor_sync(0xffffffffu, rmin[q], o));
            rmax[q] = fmax(rmax[q], __shfl_xor_sync(0xffffffffu, rmax[q], o));
        }
        if (j < d && lane == 0) {
            mypart[off_reg + j] = rmin[q];
            mypart[off_reg + d + j] = rmax[q];
        }
    }
    double* tot = sm;
    const bool last = hier_reduce(A.ws.part, E, A.ws.gpart, A.ws.tickets, tot,
                                  [=](int e) { return e < off_reg ? 0 : (e < off_reg + d ? 1 : 2); });
    if (!last) return;
    for (int t = tid; t < 2 * d; t += kTlThreads) A.ws.reg[t] = tot[off_reg + t];
    if (!lin) return;
    if (A.world > 1) peer_block_allreduce(PD, tot, off_reg, 0);
    for (int t = tid; t < off_reg; t += kTlThreads) A.ws.vecA[t] = tot[t];
    double* Ls = sm + ((E + 3) & ~3);
    double* diag0 = Ls + M * M;
    double* Z = diag0 + M;
    for (int t = tid; t < M * M; t += kTlThreads) Ls[t] = tot[t];
    __syncthreads();
    chol_factor(Ls, M, diag0, &bad);
    for (int t = tid; t < M * M; t += kTlThreads) A.ws.L[t] = Ls[t];
    chol_solve(Ls, M, tot + offY, P, Z, A.ws.beta, bad);
    if (tid == 0 && bad) atomicOr(A.status, ABCDLS_S_SINGULAR);
}

// beta | gamma in shared memory with an even row pitch Pp so that two neighbouring parameters are one 16-byte load
__device__ __forceinline__ void load_coef(double* dst, const double* __restrict__ src, int M, int P, int Pp) {
    for (int t = threadIdx.x; t < M * Pp; t += blockDim.x) {
        const int a = t / Pp, p = t - a * Pp;
        dst[t] = p < P ? src[a * P + p] : 0.0;
    }
}

// the_m = colMeans(residuals) from the column sums of T1 (header comment); identical in every CTA
__device__ __forceinline__ double the_m_of(const double* usum, const double* sbeta, int d, int Pp, int p) {
    const int M = d + 1;
    const double n = usum[0];
    double m = usum[M + p] / n - sbeta[p];
    for (int j = 0; j < d; ++j) m = fma(-sbeta[(j + 1) * Pp + p], usum[1 + j] / n, m);
    return m;
}

// =====================================================================================================================
// T2 (hcorr only): X'W log(res^2).  Thread (row rr, half h) keeps the row's statistics in registers (DC >= d) and
// computes the residuals of parameters {4k + 2h, 4k + 2h + 1}; the tile [ones | xs | log res^2 | w] feeds the DMMA.
template <int DC>
__global__ void __launch_bounds__(kTlThreads, DC > 16 ? 1 : 2) k_tail_resid_fit(TailArgs A, PeerDev PD) {
    pdl_prologue();
    extern __shared__ __align__(16) double sm[];
    const int d = A.d, P = A.P, M = d + 1, W = M + P, Pp = (P + 3) & ~3;
    const int tid = threadIdx.x;
    double* T = sm;
    double* red = sm;
    double* sbeta = sm + region0_doubles(d, P);             // [M][Pp]
    double* s_m = sbeta + M * Pp;                           // [Pp]
    const int offU = M * M + M * P;
    load_coef(sbeta, A.ws.beta, M, P, Pp);
    for (int t = tid; t < kTlPitch; t += kTlThreads) {
        T[t] = 1.0;
        T[(size_t)(W + 1) * kTlPitch + t] = 0.0;
    }
    __syncthreads();
    if (tid < Pp) s_m[tid] = tid < P ? the_m_of(A.ws.vecA + offU, sbeta, d, Pp, tid) : 0.0;
    long long n_acc = *A.n_acc;
    if (n_acc > A.cap) n_acc = A.cap;
    const long long ntile = (n_acc + kTlRows - 1) / kTlRows;
    const long long per = (ntile + gridDim.x - 1) / gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per, t_end = min(ntile, t_begin + per);
    constexpr int NA = DC <= 2 ? 1 : (DC <= 16 ? 3 : 5);      // d <= DC  =>  M = d + 1 rows fit NA tiles of 8
    const MmaMap<NA> mm = make_mma_map<NA>(M, W, M, false);
    double C[NA * kMB][2];
#pragma unroll
    for (int t = 0; t < NA * kMB; ++t) C[t][0] = C[t][1] = 0.0;
    const int rr = tid & (kTlRows - 1), h = tid >> 7;
    __syncthreads();

    for (long long tile = t_begin; tile < t_end; ++tile) {
        const long long r0 = tile * kTlRows;
        const int nr = (int)min((long long)kTlRows, n_acc - r0);
        const bool valid = rr < nr;
        const size_t r = (size_t)(r0 + rr);
        double x[DC];
#pragma unroll
        for (int j = 0; j < DC; ++j) x[j] = (valid && j < d) ? A.xs[(size_t)j * A.cap + r] : 0.0;
        if (h == 0) {
#pragma unroll
            for (int j = 0; j < DC; ++j)
                if (j < d) T[(size_t)(1 + j) * kTlPitch + rr] = x[j];
            T[(size_t)W * kTlPitch + rr] = valid ? A.w[r] : 0.0;
        }
        for (int p = 4 * h; p < P; p += 8) {                 // four parameters per step: four independent FMA chains
            double yv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) yv[k] = (valid && p + k < P) ? A.y[(size_t)(p + k) * A.cap + r] : 0.0;
            double2 f0 = *reinterpret_cast<const double2*>(sbeta + p);
            double2 f1 = *reinterpret_cast<const double2*>(sbeta + p + 2);
#pragma unroll
            for (int j = 0; j < DC; ++j)
                if (j < d) {
                    const double2 b0 = *reinterpret_cast<const double2*>(sbeta + (j + 1) * Pp + p);
                    const double2 b1 = *reinterpret_cast<const double2*>(sbeta + (j + 1) * Pp + p + 2);
                    f0.x = fma(x[j], b0.x, f0.x);
                    f0.y = fma(x[j], b0.y, f0.y);
                    f1.x = fma(x[j], b1.x, f1.x);
                    f1.y = fma(x[j], b1.y, f1.y);
                }
            const double fv[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p + k < P) {
                    const double v = (yv[k] - fv[k]) - s_m[p + k];
                    T[(size_t)(M + p + k) * kTlPitch + rr] = valid ? log(v * v) : 0.0;
                }
        }
        __syncthreads();
        mma_accumulate<NA>(T, W, mm, C);
        __syncthreads();
    }
    const int E = M * P;
    double* mypart = A.ws.part + (size_t)blockIdx.x * E;
    mma_combine<NA>(red, mm, C, M, W, P, 0, 0, mypart);
    double* tot = sm;
    const bool last = hier_reduce(A.ws.part, E, A.ws.gpart, A.ws.tickets + (A.ws.maxgroups + 1), tot,
                                  [](int) { return 0; });
    if (!last) return;
    if (A.world > 1) peer_block_allreduce(PD, tot, E, 0);
    double* Ls = sm + ((E + 3) & ~3);
    double* Z = Ls + M * M;
    for (int t = tid; t < M * M; t += kTlThreads) Ls[t] = A.ws.L[t];
    __syncthreads();
    chol_solve(Ls, M, tot, P, Z, A.ws.gamma, 0);
    // R: lsfit(..., log(residuals^2)) stops on a non-finite response (a residual that is exactly zero)
    int nonfin = 0;
    for (int t = tid; t < M * P; t += kTlThreads)
        if (!isfinite(A.ws.gamma[t])) nonfin = 1;
    if (__syncthreads_or(nonfin) && tid == 0 && !(*A.status & ABCDLS_S_SINGULAR)) atomicOr(A.status, ABCDLS_S_NONFINITE_FIT);
}

// =====================================================================================================================
// T3: one lane per accepted row, the row's statistics in registers (DC >= d; DC == 0: rejection, the sample is y).
// Column statistics without shuffles: the warp parks its 32 x P values in shared memory and lane p walks column p.
constexpr int kT3Stat = 7;   // min, max, sum x, sum w x, sum w x^2, sum w, sum w res^2 (uncorrected residual: sigma2)
template <int DC>
__global__ void __launch_bounds__(kTlThreads, 2) k_tail_adjust_final(TailArgs A, PeerDev PD) {
    pdl_prologue();
    extern __shared__ __align__(16) double sm[];
    const int d = A.d, P = A.P, M = d + 1, Pp = (P + 3) & ~3;
    constexpr bool lin = DC > 0;
    const bool hcorr = lin && A.hcorr;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* park = sm + (size_t)wid * (kTlMaxP * 33 + 32);   // [P][33] values | [32] weights, per warp
    double* sbeta = sm + (size_t)kTlWarps * (kTlMaxP * 33 + 32);   // [M][Pp]
    double* sgam = sbeta + M * Pp;                           // [M][Pp]
    double* s_m = sgam + M * Pp;                             // [Pp]
    double* tsc = s_m + Pp;                                  // [d]  scaled target (final)
    const int offU = M * M + M * P;
    if (lin) {
        load_coef(sbeta, A.ws.beta, M, P, Pp);
        if (hcorr) load_coef(sgam, A.ws.gamma, M, P, Pp);
        else
            for (int t = tid; t < M * Pp; t += kTlThreads) sgam[t] = 0.0;
        __syncthreads();
        if (tid < Pp) s_m[tid] = tid < P ? the_m_of(A.ws.vecA + offU, sbeta, d, Pp, tid) : 0.0;
    }
    long long n_acc = *A.n_acc;
    if (n_acc > A.cap) n_acc = A.cap;
    const long long nbatch = (n_acc + 31) / 32;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double a[kT3Stat] = {inf, -inf, 0.0, 0.0, 0.0, 0.0, 0.0};     // lane p: statistics of parameter p (this warp's rows)
    __syncthreads();

    for (long long bt = (long long)blockIdx.x * kTlWarps + wid; bt < nbatch; bt += (long long)gridDim.x * kTlWarps) {
        const long long r = bt * 32 + lane;
        const int nv = (int)min(32ll, n_acc - bt * 32);
        const bool valid = lane < nv;
        double x[DC > 0 ? DC : 1];
        double wv = valid ? 1.0 : 0.0;
        if (lin) {
#pragma unroll
            for (int j = 0; j < DC; ++j) x[j] = (valid && j < d) ? A.xs[(size_t)j * A.cap + r] : 0.0;
            if (valid) wv = A.w[r];
        }
        park[kTlMaxP * 33 + lane] = wv;
        // the posterior sample (adj, or y for rejection) is parked per warp; sum w res^2 uses the uncorrected residual
        for (int p = 0; p < P; p += 4) {                     // four parameters per step: eight independent FMA chains
            double yv[4], ov[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                yv[k] = (valid && p + k < P) ? A.y[(size_t)(p + k) * A.cap + r] : 0.0;
                ov[k] = yv[k];
            }
            if (lin) {
                double2 f0 = *reinterpret_cast<const double2*>(sbeta + p);
                double2 f1 = *reinterpret_cast<const double2*>(sbeta + p + 2);
                double2 g0 = make_double2(0.0, 0.0), g1 = make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < DC; ++j)
                    if (j < d) {
                        const double2 b0 = *reinterpret_cast<const double2*>(sbeta + (j + 1) * Pp + p);
                        const double2 b1 = *reinterpret_cast<const double2*>(sbeta + (j + 1) * Pp + p + 2);
                        const double2 c0 = *reinterpret_cast<const double2*>(sgam + (j + 1) * Pp + p);
                        const double2 c1 = *reinterpret_cast<const double2*>(sgam + (j + 1) * Pp + p + 2);
                        f0.x = fma(x[j], b0.x, f0.x);
                        f0.y = fma(x[j], b0.y, f0.y);
                        f1.x = fma(x[j], b1.x, f1.x);
                        f1.y = fma(x[j], b1.y, f1.y);
                        g0.x = fma(x[j], c0.x, g0.x);
                        g0.y = fma(x[j], c0.y, g0.y);
                        g1.x = fma(x[j], c1.x, g1.x);
                        g1.y = fma(x[j], c1.y, g1.y);
                    }
                const double fv[4] = {f0.x, f0.y, f1.x, f1.y}, gv[4] = {g0.x, g0.y, g1.x, g1.y};
                double qv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double v = (yv[k] - fv[k]) - s_m[p + k];
                    qv[k] = wv * v * v;                      // sigma2 uses the uncorrected residual
                    // pred.sd * res / pred.si = res * sqrt(exp(g0)) / sqrt(exp(g0 + g)) = res * exp(-g / 2): one exp
                    if (hcorr) v *= exp(-0.5 * gv[k]);
                    ov[k] = (sbeta[p + k] + s_m[p + k]) + v;
                    if (valid && p + k < P) {
                        A.res[(size_t)(p + k) * A.cap + r] = v;
                        A.adj[(size_t)(p + k) * A.cap + r] = ov[k];
                    }
                }
                // sum w res^2: one shuffle tree for the four parameters, kept by the lane that owns the parameter
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int k = 0; k < 4; ++k) qv[k] += __shfl_xor_sync(0xffffffffu, qv[k], o);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (lane == p + k) a[6] += qv[k];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p + k < P) park[(size_t)(p + k) * 33 + lane] = ov[k];
        }
        __syncwarp();
        if (lane < P) {
            const double* col = park + (size_t)lane * 33;
            const double* wt = park + kTlMaxP * 33;
            for (int i = 0; i < nv; ++i) {
                const double xv = col[i], ww = wt[i];
                a[0] = fmin(a[0], xv);
                a[1] = fmax(a[1], xv);
                a[2] += xv;
                a[3] = fma(ww, xv, a[3]);
                a[4] = fma(ww * xv, xv, a[4]);
                a[5] += ww;
            }
        }
        __syncwarp();
    }
    // ---- CTA partial [P][7]: warps combined in order ----
    __syncthreads();
    double* segs = sm;                                       // [warps][P][7]
    if (lane < P)
        for (int q = 0; q < kT3Stat; ++q) segs[((size_t)wid * P + lane) * kT3Stat + q] = a[q];
    __syncthreads();
    const int E = kT3Stat * P;
    double* mypart = A.ws.part + (size_t)blockIdx.x * E;
    for (int e = tid; e < E; e += kTlThreads) {
        const int q = e % kT3Stat;
        double acc = segs[e];
        for (int s = 1; s < kTlWarps; ++s) {
            const double v = segs[(size_t)s * E + e];
            acc = q == 0 ? fmin(acc, v) : (q == 1 ? fmax(acc, v) : acc + v);
        }
        mypart[e] = acc;
    }
    __syncthreads();
    double* stats = sm;                                      // [P][7] (this rank)
    const bool last = hier_reduce(A.ws.part, E, A.ws.gpart, A.ws.tickets + 2 * (A.ws.maxgroups + 1), stats,
                                  [](int e) { return e % kT3Stat == 0 ? 1 : (e % kT3Stat == 1 ? 2 : 0); });
    if (!last) return;

    // ---- end of the call: ONE exchange for everything that is a min, a max (negated), an OR (negated bits) or a sum,
    //      then the packed small-result buffer (layout: loclinear.cu:k_final_small) ----
    double* vec = sm + ((E + 3) & ~3);
    const int n_min = 2 * d + 2 * P + 8, count = n_min + 5 * P;
    int st = *A.status;
    if (A.world > 1 && ld_volatile_u64(PD.err) != 0ull) st |= ABCDLS_S_PEER_TIMEOUT;
    for (int t = tid; t < count; t += kTlThreads) {
        double v;
        if (t < d) v = A.ws.reg[t];
        else if (t < 2 * d) v = -A.ws.reg[t];
        else if (t < 2 * d + P) v = stats[(t - 2 * d) * kT3Stat];
        else if (t < 2 * d + 2 * P) v = -stats[(t - 2 * d - P) * kT3Stat + 1];
        else if (t < n_min) v = -(double)((st >> (t - 2 * d - 2 * P)) & 1);
        else v = stats[((t - n_min) / 5) * kT3Stat + 2 + (t - n_min) % 5];
        vec[t] = v;
    }
    __syncthreads();
    if (A.world > 1) {
        peer_block_allreduce(PD, vec, count, n_min);
        if (tid == 0 && ld_volatile_u64(PD.err) != 0ull) vec[2 * d + 2 * P + 6] = -1.0;    // bit 6: PEER_TIMEOUT
        __syncthreads();
    }
    __shared__ int flags[2];
    __shared__ double lgs[kTlMaxP];
    double* small = A.small_out;
    double* host = small;
    double* o_mad = small + 8;
    double* o_stats = o_mad + d;
    double* o_sig = o_stats + 6 * P;
    double* o_pred = o_sig + P;
    double* o_coef = o_pred + P;
    double* o_coefh = o_coef + M * P;
    if (tid < 2) flags[tid] = 1;
    __syncthreads();
    if (tid < d) {
        const double m = A.mad[tid];
        o_mad[tid] = m;
        tsc[tid] = (m == 0.0) ? A.target[tid] : __ddiv_rn(A.target[tid], m);
        if (!(vec[tid] == -vec[d + tid])) atomicAnd(&flags[0], 0);        // some accepted statistic varies
        if (m != 0.0) atomicAnd(&flags[1], 0);
    }
    if (tid < P) {
        o_stats[tid * 6] = vec[2 * d + tid];
        o_stats[tid * 6 + 1] = -vec[2 * d + P + tid];
#pragma unroll
        for (int k = 0; k < 4; ++k) o_stats[tid * 6 + 2 + k] = vec[n_min + tid * 5 + k];
    }
    __syncthreads();
    if (lin && tid < P) {
        const double s2 = vec[n_min + tid * 5 + 4] / A.ws.vecA[0];       // sum w res^2 / sum w (both over all ranks)
        o_sig[tid] = s2;
        lgs[tid] = log(s2);
        o_pred[tid] = sbeta[tid] + s_m[tid];
        double acc = 0.0, acch = 0.0;
        for (int j = 0; j < d; ++j) {
            const double b = sbeta[(j + 1) * Pp + tid];
            o_coef[(j + 1) * P + tid] = b;
            acc = __dadd_rn(acc, __dmul_rn(tsc[j], b));
            if (hcorr) {
                const double g = sgam[(j + 1) * Pp + tid];
                o_coefh[(j + 1) * P + tid] = g;
                acch = __dadd_rn(acch, __dmul_rn(tsc[j], g));
            }
        }
        o_coef[tid] = sbeta[tid] - acc;
        if (hcorr) o_coefh[tid] = sgam[tid] - acch;
    }
    __syncthreads();
    if (tid == 0) {
        int so = 0;
        for (int b = 0; b < 8; ++b)
            if (vec[2 * d + 2 * P + b] != 0.0) so |= 1 << b;
        double ls = 0.0;
        if (lin)
            for (int p = 0; p < P; ++p) ls += lgs[p];
        host[0] = (double)so;
        host[1] = (double)*A.n_acc;
        host[2] = *A.ds;
        host[3] = (double)flags[0];
        host[4] = (double)flags[1];
        host[5] = ls;
        host[6] = host[7] = 0.0;
    }
}

// =====================================================================================================================
// postpr (rejection): the end of the call in one block - the per-model counts of all ranks and the OR of the status
// words in ONE exchange, then the packed small buffer  [status, n_acc, ds, every MAD zero?, 0, 0, 0, 0 | mad (d) | counts (K)].
__global__ void __launch_bounds__(256) k_postpr_final(PeerDev PD, int world, const int* __restrict__ status,
                                                      const long long* __restrict__ n_acc, const double* __restrict__ ds,
                                                      const double* __restrict__ mad, int d,
                                                      const long long* __restrict__ counts, int K,
                                                      double* __restrict__ small) {
    pdl_prologue();
    extern __shared__ double vec[];   // [8 status bits (negated: min) | K counts (sum)]
    const int tid = threadIdx.x;
    int st = *status;
    if (world > 1 && ld_volatile_u64(PD.err) != 0ull) st |= ABCDLS_S_PEER_TIMEOUT;
    for (int t = tid; t < 8 + K; t += blockDim.x) vec[t] = t < 8 ? -(double)((st >> t) & 1) : (double)counts[t - 8];
    __syncthreads();
    if (world > 1) {
        peer_block_allreduce(PD, vec, 8 + K, 8);
        if (tid == 0 && ld_volatile_u64(PD.err) != 0ull) vec[6] = -1.0;
        __syncthreads();
    }
    __shared__ int allzero;
    if (tid == 0) allzero = 1;
    __syncthreads();
    for (int j = tid; j < d; j += blockDim.x) {
        small[8 + j] = mad[j];
        if (mad[j] != 0.0) atomicAnd(&allzero, 0);
    }
    for (int k = tid; k < K; k += blockDim.x) small[8 + d + k] = vec[8 + k];
    __syncthreads();
    if (tid == 0) {
        int so = 0;
        for (int b = 0; b < 8; ++b)
            if (vec[b] != 0.0) so |= 1 << b;
        small[0] = (double)so;
        small[1] = (double)*n_acc;
        small[2] = *ds;
        small[3] = (double)allzero;
        small[4] = small[5] = small[6] = small[7] = 0.0;
    }
}

// What a caller normally reads from a result (small buffer, accepted rows, weights, posterior sample) leaves the
// buffers of the replayed CUDA graph in ONE launch: up to four 2-D segments (rows x cols, source pitch ld) are
// packed back to back into a freshly allocated block.  (Four framework clone calls cost ~35 us of host time per
// call - the GPU idles meanwhile, because the next call cannot be enqueued before this one's results are safe.)
struct CopySegs {
    const double* src[4];
    long long ld[4], rows[4], cols[4], off[4];
    int n;
};
constexpr int kCopyChunk = 4096;
__global__ void __launch_bounds__(256) k_copy_out(CopySegs S, double* __restrict__ dst) {
    const int sgi = blockIdx.y;
    if (sgi >= S.n) return;
    const long long cols = S.cols[sgi], nch = (cols + kCopyChunk - 1) / kCopyChunk;
    for (long long b = blockIdx.x; b < S.rows[sgi] * nch; b += gridDim.x) {
        const long long r = b / nch, c0 = (b - r * nch) * kCopyChunk;
        const double* __restrict__ src = S.src[sgi] + r * S.ld[sgi];
        double* __restrict__ out = dst + S.off[sgi] + r * cols;
        const long long c1 = min(cols, c0 + (long long)kCopyChunk);
        for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) out[c] = src[c];
    }
}

size_t tail_E(int d, int P) {
    const int M = d + 1, W = M + P;
    return (size_t)M * M + (size_t)M * P + W + 2 * d;
}

int tail_grid(abcdls_handle_t h, long long units, int per_sm) {   // units: 128-row tiles (T1, T2) / 256-row batches (T3)
    long long b = units;
    const long long mx = (long long)h->sm_count * per_sm;
    if (b > mx) b = mx;
    if (b < 1) b = 1;
    return (int)b;
}

size_t tail_region0(int d, int P) {
    const size_t tile = (size_t)(d + 1 + P + 2 + d) * kTlPitch;
    const size_t red = (size_t)kTlWarps * kMT * 64;
    return std::max(tile, red);
}
size_t tail_smem1(int d, int P) { return (tail_region0(d, P) + 3 * (size_t)d + 8) * sizeof(double); }
size_t tail_smem2(int d, int P) {
    const int M = d + 1, Pp = (P + 3) & ~3;
    return (tail_region0(d, P) + (size_t)M * Pp + Pp) * sizeof(double);
}
size_t tail_smem3(int d, int P) {
    const int M = d + 1, Pp = (P + 3) & ~3;
    return ((size_t)kTlWarps * (kTlMaxP * 33 + 32) + 2 * (size_t)M * Pp + Pp + d) * sizeof(double);
}

TailWs carve_tail(void* ws, int d, int P, int sm_count) {
    TailWs w;
    const int M = d + 1;
    const size_t E = tail_E(d, P);
    const int G = sm_count * 3;
    w.maxgroups = (G + kTlGroup - 1) / kTlGroup;
    char* p = reinterpret_cast<char*>(ws);
    auto take = [&](size_t bytes) {
        char* q = p;
        p += align_up(bytes, 256);
        return q;
    };
    w.tickets = reinterpret_cast<unsigned*>(take((size_t)3 * (w.maxgroups + 1) * sizeof(unsigned)));
    w.part = reinterpret_cast<double*>(take((size_t)G * E * 8));
    w.gpart = reinterpret_cast<double*>(take((size_t)w.maxgroups * E * 8));
    w.vecA = reinterpret_cast<double*>(take(E * 8));
    w.reg = reinterpret_cast<double*>(take((size_t)2 * d * 8));
    w.L = reinterpret_cast<double*>(take((size_t)M * M * 8));
    w.beta = reinterpret_cast<double*>(take((size_t)M * P * 8));
    w.gamma = reinterpret_cast<double*>(take((size_t)M * P * 8));
    return w;
}

size_t tail_ws_bytes(int d, int P, int sm_count) {
    TailWs w = carve_tail(nullptr, d, P, sm_count);
    return (size_t)(reinterpret_cast<char*>(w.gamma) - reinterpret_cast<char*>(0)) +
           align_up((size_t)(d + 1) * P * 8, 256);
}

// kernels ask for more than 48 KB of dynamic shared memory: raise the limit once per kernel (and when a call needs more)
template <class K>
int ensure_smem(abcdls_handle_t h, K kernel, size_t smem, size_t* have) {
    if (smem > *have) {
        ABCDLS_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        *have = smem;
    }
    return ABCDLS_OK;
}

int dc_index(int d) { return d <= 2 ? 0 : (d <= 8 ? 1 : (d <= 16 ? 2 : 3)); }

}  // namespace

extern "C" {

int abcdls_tail_supported(int d, int P) { return d >= 1 && d <= kTlMaxD && P >= 1 && P <= kTlMaxP; }

size_t abcdls_tail_workspace_bytes(abcdls_handle_t h, int d, int P) {
    if (!h || !abcdls_tail_supported(d, P)) return 0;
    return tail_ws_bytes(d, P, h->sm_count);
}

// ws: abcdls_tail_workspace_bytes, zeroed ONCE by the caller (the kernels leave their ticket counters zeroed).
int abcdls_tail_gather_fit(abcdls_handle_t h, abcdls_peer_t peer, const double* cand_ss, int64_t ld_cand,
                           const int64_t* slot, const int64_t* perm, const double* ss, int64_t ld_ss,
                           const double* param_cm, int64_t ld_param, const double* param_rm, int64_t row_pitch,
                           const double* dist_acc, const int64_t* idx, const int64_t* n_acc, int64_t row_offset,
                           int64_t cap, int d, int P, const double* mad, const double* target, const double* ds,
                           int loclinear, double* xs, double* y, double* ss_out, double* w, void* ws, size_t ws_bytes,
                           int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && dist_acc && idx && n_acc && mad && target && ds && y && ss_out && ws && status_dev,
                     "tail_gather_fit: null pointer");
    ABCDLS_CHECK_ARG(h, (cand_ss && slot) || ss, "tail_gather_fit: no source of the statistics");
    ABCDLS_CHECK_ARG(h, (param_cm != nullptr) != (param_rm != nullptr), "tail_gather_fit: one parameter table");
    ABCDLS_CHECK_ARG(h, !loclinear || (xs && w), "tail_gather_fit: xs / w");
    ABCDLS_CHECK_ARG(h, abcdls_tail_supported(d, P) && cap >= 1, "tail_gather_fit: shape");
    if (ws_bytes < tail_ws_bytes(d, P, h->sm_count)) {
        h->err = "tail workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    TailArgs A = {};
    A.cand = cand_ss; A.cand_pitch = ld_cand; A.slot = (const long long*)slot; A.perm = (const long long*)perm;
    A.ss = ss; A.ld_ss = ld_ss; A.par_cm = param_cm; A.ld_par = ld_param; A.par_rm = param_rm; A.par_pitch = row_pitch;
    A.dacc = dist_acc; A.idx = (const long long*)idx; A.n_acc = (const long long*)n_acc; A.row_offset = row_offset;
    A.cap = cap; A.mad = mad; A.target = target; A.ds = ds; A.xs = xs; A.y = y; A.ss_out = ss_out; A.w = w;
    A.d = d; A.P = P; A.lin = loclinear ? 1 : 0; A.hcorr = 0; A.world = peer ? peer->dev.world : 1;
    A.status = status_dev;
    A.ws = carve_tail(ws, d, P, h->sm_count);
    PeerDev PD = {};
    if (peer) PD = peer->dev;
    const size_t smem = tail_smem1(d, P);
    static size_t have[3] = {48 * 1024, 48 * 1024, 48 * 1024};
    const int rowsA = d + 2;                                 // intercept + d slopes + the unweighted ones row
    const int grid = tail_grid(h, (cap + kTlRows - 1) / kTlRows, 2);
    cudaStream_t st = (cudaStream_t)stream;                                 // intercept + d slopes + the unweighted ones row
    if (rowsA <= 8) {
        if (int rc = ensure_smem(h, k_tail_gather_fit<1>, smem, &have[0])) return rc;
        launch_pdl(k_tail_gather_fit<1>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
    } else if (rowsA <= 24) {
        if (int rc = ensure_smem(h, k_tail_gather_fit<3>, smem, &have[1])) return rc;
        launch_pdl(k_tail_gather_fit<3>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
    } else {
        if (int rc = ensure_smem(h, k_tail_gather_fit<5>, smem, &have[2])) return rc;
        launch_pdl(k_tail_gather_fit<5>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_tail_resid_fit(abcdls_handle_t h, abcdls_peer_t peer, const double* xs, const double* y, const double* w,
                          const int64_t* n_acc, int64_t cap, int d, int P, void* ws, size_t ws_bytes, int* status_dev,
                          void* stream) {
    ABCDLS_CHECK_ARG(h, h && xs && y && w && n_acc && ws && status_dev, "tail_resid_fit: null pointer");
    ABCDLS_CHECK_ARG(h, abcdls_tail_supported(d, P) && cap >= 1, "tail_resid_fit: shape");
    if (ws_bytes < tail_ws_bytes(d, P, h->sm_count)) {
        h->err = "tail workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    TailArgs A = {};
    A.xs = const_cast<double*>(xs); A.y = const_cast<double*>(y); A.w = const_cast<double*>(w);
    A.n_acc = (const long long*)n_acc; A.cap = cap; A.d = d; A.P = P; A.lin = 1; A.hcorr = 1;
    A.world = peer ? peer->dev.world : 1;
    A.status = status_dev;
    A.ws = carve_tail(ws, d, P, h->sm_count);
    PeerDev PD = {};
    if (peer) PD = peer->dev;
    const size_t smem = tail_smem2(d, P);
    static size_t have[4] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
    const int grid = tail_grid(h, (cap + kTlRows - 1) / kTlRows, d > 16 ? 1 : 2);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dc_index(d)) {
        case 0:
            if (int rc = ensure_smem(h, k_tail_resid_fit<2>, smem, &have[0])) return rc;
            launch_pdl(k_tail_resid_fit<2>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        case 1:
            if (int rc = ensure_smem(h, k_tail_resid_fit<8>, smem, &have[1])) return rc;
            launch_pdl(k_tail_resid_fit<8>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        case 2:
            if (int rc = ensure_smem(h, k_tail_resid_fit<16>, smem, &have[2])) return rc;
            launch_pdl(k_tail_resid_fit<16>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        default:
            if (int rc = ensure_smem(h, k_tail_resid_fit<32>, smem, &have[3])) return rc;
            launch_pdl(k_tail_resid_fit<32>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_tail_adjust_final(abcdls_handle_t h, abcdls_peer_t peer, const double* xs, const double* y, const double* w,
                             const int64_t* n_acc, const double* ds, int64_t cap, int d, int P, int loclinear,
                             int hcorr, const double* mad, const double* target, double* adj, double* res, void* ws,
                             size_t ws_bytes, double* small_out, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && y && n_acc && ds && mad && target && ws && small_out && status_dev,
                     "tail_adjust_final: null pointer");
    ABCDLS_CHECK_ARG(h, !loclinear || (xs && w && adj && res), "tail_adjust_final: loclinear blocks");
    ABCDLS_CHECK_ARG(h, abcdls_tail_supported(d, P) && cap >= 1, "tail_adjust_final: shape");
    if (ws_bytes < tail_ws_bytes(d, P, h->sm_count)) {
        h->err = "tail workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    TailArgs A = {};
    A.xs = const_cast<double*>(xs); A.y = const_cast<double*>(y); A.w = const_cast<double*>(w);
    A.n_acc = (const long long*)n_acc; A.ds = ds; A.cap = cap; A.d = d; A.P = P; A.lin = loclinear ? 1 : 0;
    A.hcorr = hcorr ? 1 : 0; A.mad = mad; A.target = target; A.adj = adj; A.res = res; A.small_out = small_out;
    A.world = peer ? peer->dev.world : 1;
    A.status = status_dev;
    A.ws = carve_tail(ws, d, P, h->sm_count);
    PeerDev PD = {};
    if (peer) PD = peer->dev;
    const size_t smem = tail_smem3(d, P);
    static size_t have[5] = {48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024, 48 * 1024};
    const int grid = tail_grid(h, (cap + kTlThreads - 1) / kTlThreads, 2);
    cudaStream_t st = (cudaStream_t)stream;
    switch (loclinear ? dc_index(d) : 4) {
        case 0:
            if (int rc = ensure_smem(h, k_tail_adjust_final<2>, smem, &have[0])) return rc;
            launch_pdl(k_tail_adjust_final<2>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        case 1:
            if (int rc = ensure_smem(h, k_tail_adjust_final<8>, smem, &have[1])) return rc;
            launch_pdl(k_tail_adjust_final<8>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        case 2:
            if (int rc = ensure_smem(h, k_tail_adjust_final<16>, smem, &have[2])) return rc;
            launch_pdl(k_tail_adjust_final<16>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        case 3:
            if (int rc = ensure_smem(h, k_tail_adjust_final<32>, smem, &have[3])) return rc;
            launch_pdl(k_tail_adjust_final<32>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
        default:
            if (int rc = ensure_smem(h, k_tail_adjust_final<0>, smem, &have[4])) return rc;
            launch_pdl(k_tail_adjust_final<0>, dim3(grid), dim3(kTlThreads), smem, st, A, PD);
            break;
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// nseg <= 4 segments; segment i = rows[i] x cols[i] 8-byte elements read with pitch ld[i] from src[i], written densely
// (row after row) to dst, segment after segment.
int abcdls_copy_out(abcdls_handle_t h, int nseg, const void* const* src, const int64_t* ld, const int64_t* rows,
                    const int64_t* cols, void* dst, void* stream) {
    ABCDLS_CHECK_ARG(h, h && src && ld && rows && cols && dst && nseg >= 1 && nseg <= 4, "copy_out");
    CopySegs S = {};
    S.n = nseg;
    long long off = 0, most = 1;
    for (int i = 0; i < nseg; ++i) {
        ABCDLS_CHECK_ARG(h, (src[i] || rows[i] * cols[i] == 0) && rows[i] >= 0 && cols[i] >= 0 && ld[i] >= cols[i], "copy_out: segment");
        S.src[i] = reinterpret_cast<const double*>(src[i]);
        S.ld[i] = ld[i];
        S.rows[i] = rows[i];
        S.cols[i] = cols[i];
        S.off[i] = off;
        off += rows[i] * cols[i];
        const long long blocks = rows[i] * ((cols[i] + kCopyChunk - 1) / kCopyChunk);
        if (blocks > most) most = blocks;
    }
    if (off == 0) return ABCDLS_OK;
    const long long mx = (long long)h->sm_count * 8;
    k_copy_out<<<dim3((unsigned)(most > mx ? mx : most), (unsigned)nseg), 256, 0, (cudaStream_t)stream>>>(
        S, reinterpret_cast<double*>(dst));
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_postpr_final(abcdls_handle_t h, abcdls_peer_t peer, const int* status_dev, const int64_t* n_acc,
                        const double* ds, const double* mad, int d, const int64_t* counts, int K, double* small_out,
                        void* stream) {
    ABCDLS_CHECK_ARG(h, h && status_dev && n_acc && ds && mad && counts && small_out, "postpr_final: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= 64 && K >= 1 && K <= 4096, "postpr_final: shape");
    PeerDev PD = {};
    if (peer) PD = peer->dev;
    launch_pdl(k_postpr_final, dim3(1), dim3(256), (size_t)(8 + K) * sizeof(double), (cudaStream_t)stream, 
        PD, peer ? peer->dev.world : 1, status_dev, (const long long*)n_acc, ds, mad, d, (const long long*)counts, K,
        small_out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
