// Batched per-column ABC on small tables: ONE CTA runs one whole abc() problem with d = P = 1 out of shared memory.
//
// ABC-DLS's own sizes are small (test_size = 10 000 rows) and its per-column mode turns every call into P separate 1-D
// problems (ABC.py:1949-1953, 2805-2809); cv4abc multiplies that by nval = 100 leave-one-out fits per column
// (ABC.py:1900-1902).  At 10^4 rows the streaming path is pure launch latency (~70 launches per fit), so here a problem
// is one thread block: the column (minus the held-out row) is loaded into shared memory once and every step of R's
// abc() runs on it in place - radix select of the median, of the MAD, of ds (order-preserving 64-bit keys, 11-bit
// digits, shared-memory histograms), the cap rule as an ordered block compaction, the 2 x 2 weighted least squares of
// loclinear with the hcorr refit, and the point estimates of summary.abc (median / mean / min / max) from a bitonic sort
// of the accepted values.  A whole cross-validation (columns x held-out rows) is one launch.
//
// Arithmetic follows the streaming path (and the oracle) statement by statement: IEEE division by the MAD, no FMA
// contraction in the distance, R's even-n median, 1.4826, ceiling(N tol) computed by the caller.
#include "common.cuh"

namespace {

constexpr int kSmT = 512;         // threads per problem: 2 CTAs per SM (48 registers, ~92 KB of shared memory each)
constexpr int kSmBins = 2048;     // 11-bit digits

struct SmallShared {
    double red[33];
    unsigned long long redu[33];
    int scan[34];
    unsigned long long prefix;
    long long krem;
    int flag;
};

__device__ __forceinline__ double blk_sum(double v, SmallShared& S) {       // fixed order: deterministic
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) S.red[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kSmT / 32; ++w) t += S.red[w];
        S.red[32] = t;
    }
    __syncthreads();
    const double r = S.red[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ unsigned long long blk_min_u64(unsigned long long v, SmallShared& S) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_down_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    if (lane == 0) S.redu[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = S.redu[0];
        for (int w = 1; w < kSmT / 32; ++w) t = S.redu[w] < t ? S.redu[w] : t;
        S.redu[32] = t;
    }
    __syncthreads();
    const unsigned long long r = S.redu[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ unsigned long long blk_max_u64(unsigned long long v, SmallShared& S) {
    return ~blk_min_u64(~v, S);
}

// k-th smallest (0-based) of the n keys keyfn(0..n-1): MSB-first radix select, digits of 11, 11, 11, 11, 11, 9 bits
template <class F>
__device__ unsigned long long blk_select(F keyfn, int n, long long k, unsigned* hist, SmallShared& S) {
    if (threadIdx.x == 0) {
        S.prefix = 0ull;
        S.krem = k;
    }
    int done = 0;
    for (int pass = 0; pass < 6; ++pass) {
        const int nb = pass < 5 ? 11 : 9;
        const int shift = 64 - done - nb;
        for (int i = threadIdx.x; i < kSmBins; i += kSmT) hist[i] = 0u;
        __syncthreads();
        const unsigned long long prefix = S.prefix;
        for (int j = threadIdx.x; j < n; j += kSmT) {
            const unsigned long long key = keyfn(j);
            if (done == 0 || (key >> (64 - done)) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & ((1u << nb) - 1u)], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {      // warp 0: every lane owns 64 consecutive bins
            const int lane = threadIdx.x;
            long long mine = 0;
            for (int i = 0; i < kSmBins / 32; ++i) mine += hist[lane * (kSmBins / 32) + i];
            long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const long long excl = incl - mine, kr = S.krem;
            if (kr >= excl && kr < incl) {      // exactly one lane
                long long run = excl;
                int bin = lane * (kSmBins / 32);
                for (;; ++bin) {
                    const long long c = hist[bin];
                    if (kr < run + c) break;
                    run += c;
                }
                S.prefix = (prefix << nb) | (unsigned long long)bin;
                S.krem = kr - run;
            }
        }
        __syncthreads();
        done += nb;
    }
    const unsigned long long r = S.prefix;
    __syncthreads();
    return r;
}

// R median of the n values keyfn encodes: order statistics (n - 1) / 2 and n / 2, mean of the two for even n
template <class F>
__device__ double blk_median(F keyfn, int n, unsigned* hist, SmallShared& S) {
    const long long k0 = (n - 1) / 2, k1 = n / 2;
    const unsigned long long v0 = blk_select(keyfn, n, k0, hist, S);
    if (k1 == k0) return keyd(v0);
    // the next order statistic: v0 again if more than k1 keys are <= v0, else the smallest key above v0
    double cle = 0.0;
    unsigned long long mgt = ~0ull;
    for (int j = threadIdx.x; j < n; j += kSmT) {
        const unsigned long long key = keyfn(j);
        if (key <= v0) cle += 1.0;
        else if (key < mgt) mgt = key;
    }
    const double c = blk_sum(cle, S);          // counts < 2^53: exact in a double
    const unsigned long long m = blk_min_u64(mgt, S);
    const unsigned long long v1 = ((long long)c > k1) ? v0 : m;
    return __dadd_rn(__dmul_rn(0.5, keyd(v0)), __dmul_rn(0.5, keyd(v1)));
}

__device__ __forceinline__ double small_dist(double x, double mad, double tsc) {
    const double a = (mad == 0.0) ? x : __ddiv_rn(x, mad);
    const double df = __dsub_rn(a, tsc);
    return __dsqrt_rn(__dmul_rn(df, df));
}

// in-place bitonic sort of (key, payload) pairs in shared memory, m = power of two
__device__ void blk_bitonic(double* key, double* pay, int m) {
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < m / 2; t += kSmT) {
                const int lo = 2 * t - (t & (stride - 1));      // index with bit `stride` clear
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const double a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                    const double pa = pay[lo];
                    pay[lo] = pay[hi];
                    pay[hi] = pa;
                }
            }
        }
    }
    __syncthreads();
}

// out[b][8] = {median, mean, ds, mad, min, max, n_acc, status} of the posterior sample of problem b
__global__ void __launch_bounds__(kSmT) k_small_columns(const double* __restrict__ ss, long long ld_ss,
                                                        const double* __restrict__ par, long long par_cs,
                                                        long long par_rs, int n_full, const int* __restrict__ col,
                                                        const long long* __restrict__ hold,
                                                        const double* __restrict__ target, long long nacc, int npad,
                                                        int loclinear, int hcorr, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char small_raw[];
    __shared__ SmallShared S;
    const int b = blockIdx.x, c = col[b];
    const long long s = hold ? hold[b] : -1;
    const int n = n_full - (s >= 0 ? 1 : 0);
    double* x = reinterpret_cast<double*>(small_raw);                         // [n_full]
    unsigned* hist = reinterpret_cast<unsigned*>(x + n_full);                  // [kSmBins]
    double* ax = reinterpret_cast<double*>(hist + kSmBins);                    // [npad] scaled statistic - scaled target
    double* ay = ax + npad;                                                    // [npad] parameter, then residual
    double* aw = ay + npad;                                                    // [npad] weight
    double* av = aw + npad;                                                    // [npad] posterior value (sorted at the end)
    const double* colp = ss + (size_t)c * ld_ss;
    const double tgt = (s >= 0) ? colp[s] : target[b];
    int status = 0;

    // ---- load the column without the held-out row; non-finite check; R cond1 (column constant) ----------------------
    unsigned long long kmin = ~0ull, kmax = 0ull;
    int nonfin = 0;
    for (int j = threadIdx.x; j < n; j += kSmT) {
        const double v = colp[j + ((s >= 0 && j >= s) ? 1 : 0)];
        x[j] = v;
        if (!isfinite(v)) nonfin = 1;
        const unsigned long long k = dkey(v);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
    if (!isfinite(tgt)) nonfin = 1;
    __syncthreads();
    kmin = blk_min_u64(kmin, S);
    kmax = blk_max_u64(kmax, S);
    if (blk_sum((double)nonfin, S) != 0.0) status |= ABCDLS_S_NONFINITE;
    if (kmin == kmax) status |= ABCDLS_S_ZERO_VAR;
    double* o = out + (size_t)b * 8;
    if (status) {       // R stops here
        if (threadIdx.x == 0) {
            for (int i = 0; i < 7; ++i) o[i] = __longlong_as_double(0x7ff8000000000000ll);
            o[7] = (double)status;
        }
        return;
    }

    // ---- normalise(): median, MAD (R: 1.4826 * median(|x - median(x)|)) --------------------------------------------
    const double med = blk_median([&](int j) { return dkey(x[j]); }, n, hist, S);
    const double mabs = blk_median([&](int j) { return dkey(fabs(__dsub_rn(x[j], med))); }, n, hist, S);
    const double mad = __dmul_rn(1.4826, mabs);
    const double tsc = (mad == 0.0) ? tgt : __ddiv_rn(tgt, mad);

    // ---- distance, ds = sort(dist)[nacc], cap rule in row order -----------------------------------------------------
    const double ds = keyd(blk_select([&](int j) { return dkey(small_dist(x[j], mad, tsc)); }, n, nacc - 1, hist, S));
    const int chunk = (n + kSmT - 1) / kSmT;
    const int j0 = min(n, threadIdx.x * chunk), j1 = min(n, j0 + chunk);
    int cnt = 0;
    for (int j = j0; j < j1; ++j) cnt += (small_dist(x[j], mad, tsc) <= ds) ? 1 : 0;
    int total = 0;
    int pos = block_exclusive_scan(cnt, S.scan, &total);
    const int n_acc = (int)min((long long)total, nacc);
    unsigned long long rmin = ~0ull, rmax = 0ull;
    for (int j = j0; j < j1; ++j) {
        const double dj = small_dist(x[j], mad, tsc);
        if (dj <= ds) {
            if (pos < n_acc) {       // wt1 & cumsum(wt1) <= nacc
                const long long row = j + ((s >= 0 && j >= s) ? 1 : 0);
                const double xv = x[j];
                const double a = (mad == 0.0) ? xv : __ddiv_rn(xv, mad);
                ax[pos] = __dsub_rn(a, tsc);
                ay[pos] = par[(size_t)c * par_cs + (size_t)row * par_rs];
                const double q = __ddiv_rn(dj, ds);
                aw[pos] = loclinear ? __dsub_rn(1.0, __dmul_rn(q, q)) : 1.0;
                const unsigned long long k = dkey(xv);
                rmin = k < rmin ? k : rmin;
                rmax = k > rmax ? k : rmax;
            }
            ++pos;
        }
    }
    __syncthreads();
    rmin = blk_min_u64(rmin, S);
    rmax = blk_max_u64(rmax, S);
    if (rmin == rmax) status |= ABCDLS_S_ZERO_VAR_REGION;       // R cond2: rejection warns, loclinear stops

    // ---- posterior sample: unadjusted parameters, or the local-linear adjustment with hcorr -------------------------
    if (!loclinear || (status & ABCDLS_S_ZERO_VAR_REGION)) {
        for (int i = threadIdx.x; i < n_acc; i += kSmT) av[i] = ay[i];
    } else {
        double s0 = 0, s1 = 0, s2 = 0, t0 = 0, t1 = 0;
        for (int i = threadIdx.x; i < n_acc; i += kSmT) {
            const double w = aw[i], xc = ax[i], y = ay[i];
            s0 += w;
            s1 += w * xc;
            s2 += w * xc * xc;
            t0 += w * y;
            t1 += w * xc * y;
        }
        s0 = blk_sum(s0, S);
        s1 = blk_sum(s1, S);
        s2 = blk_sum(s2, S);
        t0 = blk_sum(t0, S);
        t1 = blk_sum(t1, S);
        // Cholesky of [[s0, s1], [s1, s2]] with the pivot rule of k_solve
        const double dmax = fmax(fabs(s0), fabs(s2));
        const double l11sq = s0;
        const double l21 = s1 / sqrt(l11sq > 0.0 ? l11sq : 1.0);
        const double l22sq = s2 - l21 * l21;
        if (!(l11sq > 0.0) || !(l22sq > 1e-13 * s2) || !(l22sq > 0.0) || !(dmax > 0.0)) status |= ABCDLS_S_SINGULAR;
        const double l11 = sqrt(l11sq > 0.0 ? l11sq : 1.0), l22 = sqrt(l22sq > 0.0 ? l22sq : 1.0);
        auto solve2 = [&](double r0, double r1, double& b0, double& b1) {
            const double z0 = r0 / l11;
            const double z1 = (r1 - l21 * z0) / l22;
            b1 = z1 / l22;
            b0 = (z0 - l21 * b1) / l11;
        };
        double b0, b1;
        solve2(t0, t1, b0, b1);
        double rs = 0.0;
        for (int i = threadIdx.x; i < n_acc; i += kSmT) {
            const double r = ay[i] - (b0 + b1 * ax[i]);
            ay[i] = r;
            rs += r;
        }
        const double the_m = blk_sum(rs, S) / (double)n_acc;      // colMeans(residuals): unweighted
        const double pred = b0 + the_m;
        double g1 = 0.0;
        if (hcorr) {
            double u0 = 0, u1 = 0;
            for (int i = threadIdx.x; i < n_acc; i += kSmT) {
                const double r = ay[i] - the_m;
                const double z = log(r * r), w = aw[i];
                u0 += w * z;
                u1 += w * ax[i] * z;
            }
            u0 = blk_sum(u0, S);
            u1 = blk_sum(u1, S);
            double g0;
            solve2(u0, u1, g0, g1);
        }
        for (int i = threadIdx.x; i < n_acc; i += kSmT) {
            const double r = ay[i] - the_m;
            av[i] = pred + (hcorr ? r * exp(-0.5 * g1 * ax[i]) : r);      // pred.sd / pred.si = exp(-gamma1 x / 2)
        }
    }
    __syncthreads();

    // ---- summary.abc rows: min, max, (weighted) mean, (weighted) median ---------------------------------------------
    double sw = 0.0, swv = 0.0;
    unsigned long long vmin = ~0ull, vmax = 0ull;
    for (int i = threadIdx.x; i < n_acc; i += kSmT) {
        const double v = av[i], w = aw[i];
        sw += w;
        swv += w * v;
        const unsigned long long k = dkey(v);
        vmin = k < vmin ? k : vmin;
        vmax = k > vmax ? k : vmax;
    }
    sw = blk_sum(sw, S);
    swv = blk_sum(swv, S);
    vmin = blk_min_u64(vmin, S);
    vmax = blk_max_u64(vmax, S);
    for (int i = n_acc + threadIdx.x; i < npad; i += kSmT) {      // +inf padding sorts to the end
        av[i] = __longlong_as_double(0x7ff0000000000000ll);
        aw[i] = 0.0;
    }
    blk_bitonic(av, aw, npad);
    if (threadIdx.x == 0) {
        double median;
        if (loclinear && !(status & ABCDLS_S_ZERO_VAR_REGION)) {
            // quantreg::rq(x ~ 1, tau = .5, weights = w): the first sorted value whose cumulated weight reaches W / 2
            const double half = 0.5 * sw;
            double cw = 0.0;
            int kk = 0;
            for (; kk < n_acc - 1; ++kk) {
                cw += aw[kk];
                if (cw >= half) break;
            }
            median = av[kk];
        } else {                                                   // quantile(type = 7) at 0.5
            const int lo = (n_acc - 1) / 2, hi = n_acc / 2;
            median = (lo == hi || av[lo] == av[hi]) ? av[lo] : __dadd_rn(__dmul_rn(0.5, av[lo]), __dmul_rn(0.5, av[hi]));
        }
        o[0] = median;
        o[1] = swv / sw;
        o[2] = ds;
        o[3] = mad;
        o[4] = keyd(vmin);
        o[5] = keyd(vmax);
        o[6] = (double)n_acc;
        o[7] = (double)status;
    }
}

}  // namespace

static inline int small_npad(int64_t nacc) {
    int p = 2;
    while (p < nacc) p <<= 1;
    return p;
}

static inline size_t small_smem(int64_t n, int64_t nacc) {
    return (size_t)n * 8 + (size_t)kSmBins * 4 + (size_t)small_npad(nacc) * 4 * 8;
}

int abcdls_small_supported(int64_t n, int64_t nacc) {
    return n >= 3 && nacc >= 1 && nacc <= n - 1 && nacc <= 8192 && n <= (1 << 30) && small_smem(n, nacc) <= 200 * 1024;
}

int abcdls_small_columns(abcdls_handle_t h, const double* ss, int64_t ld_ss, const double* par, int64_t par_col_stride,
                         int64_t par_row_stride, int64_t n, const int32_t* col, const int64_t* hold,
                         const double* target, int B, int64_t nacc, int loclinear, int hcorr, double* out,
                         void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && par && col && out && (hold || target), "small_columns: null pointer");
    ABCDLS_CHECK_ARG(h, B >= 1 && abcdls_small_supported(n, nacc), "small_columns: table too large for one CTA");
    const size_t smem = small_smem(n, nacc);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        ABCDLS_CUDA(h, cudaFuncSetAttribute(k_small_columns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        smem_set = 200 * 1024;
    }
    k_small_columns<<<(unsigned)B, kSmT, smem, (cudaStream_t)stream>>>(ss, ld_ss, par, par_col_stride, par_row_stride, (int)n,
                                                                    col, (const long long*)hold, target, nacc,
                                                                    small_npad(nacc), loclinear, hcorr, out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}
