// summary.cu — the rows of R's summary.abc that need more than a reduction (printed by ABC-DLS on
// every abc_params call, src/Classes/ABC.py:844-861):
//   * weighted percentiles of a loclinear posterior sample - quantreg::rq(x ~ 1, tau, weights = w) is the weighted
//     order statistic "first sorted value whose cumulated weight reaches tau * sum(w)" - by an MSB radix select whose
//     histograms hold WEIGHT instead of counts (no sort of the sample);
//   * the Mode row, abc:::getmode = argmax of stats::density(x, weights = w / sum(w)) on its 512-point grid: gaussian
//     kernel, bw.nrd0, linear binning onto 2 x 512 points (C_BinDist), convolution with the kernel, linear
//     interpolation (approx) back onto [from, to].  R convolves through an FFT; the bins that can be non-zero are the
//     first 513, so the same circular convolution is a 512 x 513 direct sum here (one thread block per parameter).
// Determinism: weights are accumulated as 64-bit fixed point (2^-40 for the select, 2^-60 for the normalised density
// bins) with integer atomics - the result does not depend on the order in which threads arrive; the quantisation is
// far below the 1e-9 bar (a weighted percentile can move to a neighbouring sample value only if the cumulated weight
// hits tau * W within 1e-12 relative).
#include "common.cuh"

namespace {

constexpr int kWBins = 2048;
constexpr int kWThreads = 256;
constexpr double kWScale = 1099511627776.0;          // 2^40
constexpr double kDScale = 1152921504606846976.0;    // 2^60
constexpr int kDN = 512;                             // density(): n = 512 output points, 2 n bins

struct WJob {
    const double* x;
    unsigned long long prefix, target;
    double tau;
    int bad, pad;
};

__global__ void k_wsel_setup(WJob* jobs, int P, int ntau, const double* x, long long ld, const double* taus) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P * ntau) return;
    WJob J;
    J.x = x + (size_t)(j / ntau) * ld;
    J.prefix = 0ull;
    J.target = 0ull;
    J.tau = taus[j % ntau];
    J.bad = 0;
    J.pad = 0;
    jobs[j] = J;
}

// grid (gx, njobs): weight histogram of the current digit over the elements that match the job's prefix
__global__ void __launch_bounds__(kWThreads) k_wsel_hist(const WJob* jobs, const double* __restrict__ w, long long n,
                                                         unsigned long long* ghist, int pass) {
    __shared__ unsigned long long sh[kWBins];
    const WJob& J = jobs[blockIdx.y];
    const double* __restrict__ x = J.x;
    const unsigned long long prefix = J.prefix;
    for (int b = threadIdx.x; b < kWBins; b += kWThreads) sh[b] = 0ull;
    __syncthreads();
    const int shift_digit = pass < 5 ? 64 - 11 * (pass + 1) : 0;
    const unsigned mask = pass < 5 ? 2047u : 511u;
    const int shift_prefix = 64 - 11 * pass;
    const long long stride = (long long)gridDim.x * kWThreads;
    for (long long i = (long long)blockIdx.x * kWThreads + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = dkey(x[i]);
        if (pass == 0 || (key >> shift_prefix) == prefix) {
            const unsigned long long q = (unsigned long long)(w[i] * kWScale + 0.5);
            if (q) atomicAdd(&sh[(unsigned)(key >> shift_digit) & mask], q);
        }
    }
    __syncthreads();
    unsigned long long* g = ghist + (size_t)blockIdx.y * kWBins;
    for (int b = threadIdx.x; b < kWBins; b += kWThreads)
        if (sh[b]) atomicAdd(&g[b], sh[b]);
}

// one block per job: the bin in which the cumulated weight reaches the (remaining) target; pass 0 turns tau into a
// fixed-point target from the total weight.  Clears the histogram for the next pass.
__global__ void __launch_bounds__(kWThreads) k_wsel_pick(WJob* jobs, unsigned long long* ghist, int pass, double* out) {
    __shared__ unsigned long long part[kWThreads];
    __shared__ int found;
    WJob& J = jobs[blockIdx.x];
    unsigned long long* g = ghist + (size_t)blockIdx.x * kWBins;
    constexpr int per = kWBins / kWThreads;   // 8 consecutive bins per thread
    unsigned long long loc[per], s = 0ull;
#pragma unroll
    for (int t = 0; t < per; ++t) {
        loc[t] = g[threadIdx.x * per + t];
        s += loc[t];
    }
    part[threadIdx.x] = s;
    if (threadIdx.x == 0) found = -1;
    __syncthreads();
    if (threadIdx.x == 0) {   // serial scan of 256 partial sums: negligible, and exact
        unsigned long long run = 0ull;
        for (int t = 0; t < kWThreads; ++t) {
            const unsigned long long v = part[t];
            part[t] = run;
            run += v;
        }
        if (pass == 0) {
            double tq = J.tau * (double)run;                       // tau * sum(w) in fixed point
            unsigned long long tg = (unsigned long long)ceil(tq);
            if (tg < 1ull) tg = 1ull;
            if (tg > run) tg = run;
            J.target = tg;
            if (run == 0ull) J.bad = 1;
        }
    }
    __syncthreads();
    const unsigned long long target = J.target;
    unsigned long long run = part[threadIdx.x];
    int hit = -1;
    unsigned long long before = 0ull;
#pragma unroll
    for (int t = 0; t < per; ++t) {
        if (hit < 0 && run < target && run + loc[t] >= target) {
            hit = threadIdx.x * per + t;
            before = run;
        }
        run += loc[t];
    }
    if (hit >= 0) atomicMax(&found, hit);     // exactly one thread has a hit unless the job is bad
    __syncthreads();
    if (hit >= 0 && hit == found) {
        const int bits = pass < 5 ? 11 : 9;
        J.prefix = (J.prefix << bits) | (unsigned long long)hit;
        J.target = target - before;
        if (pass == 5) out[blockIdx.x] = keyd(J.prefix);
    }
    if (found < 0 && threadIdx.x == 0) {
        J.bad = 1;
        if (pass == 5) out[blockIdx.x] = __longlong_as_double(0x7ff8000000000000ll);
    }
    for (int b = threadIdx.x; b < kWBins; b += kWThreads) g[b] = 0ull;
}

// one block per column: mean and sum of squared deviations (two passes over the column, fixed reduction tree)
__global__ void __launch_bounds__(1024) k_col_moments(const double* __restrict__ x, long long ld, long long n,
                                                      double* __restrict__ out) {
    __shared__ double sh[32];
    __shared__ double mean_s;
    const double* col = x + (size_t)blockIdx.x * ld;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    auto block_sum = [&](double v) {
        v = warp_sum(v);
        __syncthreads();
        if (lane == 0) sh[wid] = v;
        __syncthreads();
        double t = 0.0;
        if (wid == 0) {
            t = warp_sum(sh[lane]);
        }
        return t;   // valid in warp 0
    };
    double a = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) a += col[i];
    a = block_sum(a);
    if (threadIdx.x == 0) mean_s = a / (double)n;
    __syncthreads();
    const double m = mean_s;
    double q = 0.0, r = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double dv = col[i] - m;
        q = fma(dv, dv, q);
        r += dv;
    }
    q = block_sum(q);
    __syncthreads();
    r = block_sum(r);
    if (threadIdx.x == 0) {
        out[2 * blockIdx.x] = m + r / (double)n;                       // R's mean(): one refinement pass
        out[2 * blockIdx.x + 1] = q - r * r / (double)n;               // sum (x - mean)^2
    }
}

// C_BinDist: y[ix] += w (1 - fx), y[ix + 1] += w fx with xpos = (x - lo) / xdelta, weights normalised by wnorm[p]
// (sum of the weights; NULL weights = 1 / n each).  Fixed-point integer atomics.  grid (gx, P)
__global__ void __launch_bounds__(256) k_density_bin(const double* __restrict__ x, long long ld,
                                                     const double* __restrict__ w, long long n,
                                                     const double* __restrict__ lo, const double* __restrict__ up,
                                                     const double* __restrict__ wnorm, unsigned long long* ybins) {
    __shared__ unsigned long long sh[2 * kDN];
    const int p = blockIdx.y;
    for (int b = threadIdx.x; b < 2 * kDN; b += blockDim.x) sh[b] = 0ull;
    __syncthreads();
    const double* col = x + (size_t)p * ld;
    const double l = lo[p], xdelta = (up[p] - l) / (double)(kDN - 1);   // BinDist: n grid points, 2 n bins of which the
    const double wn = w ? wnorm[p] : (double)n;                         // upper half stays empty (room for the kernel)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double xpos = (col[i] - l) / xdelta;
        double fl = floor(xpos);
        const double fx = xpos - fl;
        long long ix = (long long)fl;
        if (ix < 0) ix = 0;
        if (ix > kDN - 2) ix = kDN - 2;                                  // the data lies 7 bw inside [lo, up]
        const double wi = (w ? w[i] : 1.0) / wn;
        const unsigned long long q0 = (unsigned long long)(wi * (1.0 - fx) * kDScale + 0.5);
        const unsigned long long q1 = (unsigned long long)(wi * fx * kDScale + 0.5);
        if (q0) atomicAdd(&sh[ix], q0);
        if (q1) atomicAdd(&sh[ix + 1], q1);
    }
    __syncthreads();
    unsigned long long* g = ybins + (size_t)p * 2 * kDN;
    for (int b = threadIdx.x; b < 2 * kDN; b += blockDim.x)
        if (sh[b]) atomicAdd(&g[b], sh[b]);
}

// one block (512 threads) per column: circular convolution of the 2 n bins with the wrapped gaussian kernel (what R's
// fft(fft(y) * Conj(fft(kords)), inverse = TRUE) / length(y) computes), clamp at 0, approx() onto the 512 output
// points of [from, to], first maximum.
__global__ void __launch_bounds__(kDN) k_density_mode(const unsigned long long* __restrict__ ybins,
                                                      const double* __restrict__ lo, const double* __restrict__ up,
                                                      const double* __restrict__ frm, const double* __restrict__ to,
                                                      const double* __restrict__ bw, double* __restrict__ mode_out) {
    __shared__ double y[2 * kDN], K[2 * kDN], dens[kDN];
    __shared__ double best_v[kDN / 32];
    __shared__ int best_i[kDN / 32];
    const int p = blockIdx.x, t = threadIdx.x;
    const double l = lo[p], u = up[p], b = bw[p];
    for (int i = t; i < 2 * kDN; i += kDN) {
        y[i] = (double)ybins[(size_t)p * 2 * kDN + i] / kDScale;
        // kords <- seq(0, 2 (up - lo), length = 2 n); kords[(n + 2):(2 n)] <- kords[n:2]; dnorm(kords, sd = bw)
        const int lag = i <= kDN ? i : 2 * kDN - i;
        const double kv = (double)lag * (2.0 * (u - l)) / (double)(2 * kDN - 1);
        K[i] = exp(-0.5 * (kv / b) * (kv / b)) / (b * 2.5066282746310002);
    }
    __syncthreads();
    double s = 0.0;
    for (int j = 0; j < 2 * kDN; ++j) s = fma(y[j], K[(t - j) & (2 * kDN - 1)], s);
    dens[t] = s > 0.0 ? s : 0.0;
    __syncthreads();
    const double f = frm[p], xo = (double)t * ((to[p] - f) / (double)(kDN - 1)) + f;
    // R: approx(xords, kords, xout) with xords = seq(lo, up, length = n): the n density values sit on the binning grid
    const double gstep = (u - l) / (double)(kDN - 1);
    const double tt = (xo - l) / gstep;
    int i0 = (int)floor(tt);
    if (i0 < 0) i0 = 0;
    if (i0 > kDN - 2) i0 = kDN - 2;
    const double fr = tt - (double)i0;
    double v = dens[i0] * (1.0 - fr) + dens[i0 + 1] * fr;
    int idx = t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
    if ((t & 31) == 0) {
        best_v[t >> 5] = v;
        best_i[t >> 5] = idx;
    }
    __syncthreads();
    if (t == 0) {
        for (int q = 1; q < kDN / 32; ++q)
            if (best_v[q] > v || (best_v[q] == v && best_i[q] < idx)) {
                v = best_v[q];
                idx = best_i[q];
            }
        mode_out[p] = (double)idx * ((to[p] - f) / (double)(kDN - 1)) + f;
    }
}

}  // namespace

extern "C" {

size_t abcdls_wselect_workspace_bytes(int njobs) {
    return align_up((size_t)njobs * sizeof(WJob), 256) + (size_t)njobs * kWBins * 8;
}

// out[p * ntau + t] = weighted order statistic of column p (x: column-major, leading dimension ld, n rows; w: n weights
// >= 0) at taus_dev[t]: the smallest x whose cumulated weight (in ascending x) reaches tau * sum(w).
int abcdls_wselect(abcdls_handle_t h, const double* x, int64_t ld, const double* w, int64_t n, int P,
                   const double* taus_dev, int ntau, double* out, void* ws, size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && x && w && taus_dev && out && ws && n >= 1 && P >= 1 && ntau >= 1 && ld >= n, "wselect");
    const int njobs = P * ntau;
    if (ws_bytes < abcdls_wselect_workspace_bytes(njobs)) {
        h->err = "wselect workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    WJob* jobs = reinterpret_cast<WJob*>(ws);
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) +
                                                                    align_up((size_t)njobs * sizeof(WJob), 256));
    ABCDLS_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)njobs * kWBins * 8, st));
    k_wsel_setup<<<(njobs + 63) / 64, 64, 0, st>>>(jobs, P, ntau, x, ld, taus_dev);
    long long gx = (n + kWThreads * 8 - 1) / (kWThreads * 8);
    const long long mx = std::max(1, h->sm_count * 8 / njobs);
    if (gx > mx) gx = mx;
    for (int pass = 0; pass < 6; ++pass) {
        k_wsel_hist<<<dim3((unsigned)gx, njobs), kWThreads, 0, st>>>(jobs, w, n, hist, pass);
        k_wsel_pick<<<njobs, kWThreads, 0, st>>>(jobs, hist, pass, out);
    }
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// out[2 p] = mean(x[, p]) (R's mean: one refinement pass), out[2 p + 1] = sum (x - mean)^2
int abcdls_col_moments(abcdls_handle_t h, const double* x, int64_t ld, int64_t n, int P, double* out, void* stream) {
    ABCDLS_CHECK_ARG(h, h && x && out && n >= 1 && P >= 1 && ld >= n, "col_moments");
    k_col_moments<<<P, 1024, 0, (cudaStream_t)stream>>>(x, ld, n, out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

size_t abcdls_density_workspace_bytes(int P) { return (size_t)P * 2 * kDN * 8; }

// mode_out[p] = the grid point of R's density(x[, p], bw = bw[p], weights = w / wnorm[p], from = frm[p], to = to[p],
// n = 512) with the largest density; lo / up = from - 4 bw / to + 4 bw (density.default's binning range).
// w == NULL: equal weights.
int abcdls_density_mode(abcdls_handle_t h, const double* x, int64_t ld, const double* w, int64_t n, int P,
                        const double* lo, const double* up, const double* frm, const double* to, const double* bw,
                        const double* wnorm, double* mode_out, void* ws, size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && x && lo && up && frm && to && bw && mode_out && ws && n >= 1 && P >= 1 && ld >= n &&
                            (!w || wnorm), "density_mode");
    if (ws_bytes < abcdls_density_workspace_bytes(P)) {
        h->err = "density workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* yb = reinterpret_cast<unsigned long long*>(ws);
    ABCDLS_CUDA(h, cudaMemsetAsync(yb, 0, (size_t)P * 2 * kDN * 8, st));
    long long gx = (n + 256 * 8 - 1) / (256 * 8);
    const long long mx = std::max(1, h->sm_count * 8 / P);
    if (gx > mx) gx = mx;
    k_density_bin<<<dim3((unsigned)gx, P), 256, 0, st>>>(x, ld, w, n, lo, up, wnorm, yb);
    k_density_mode<<<P, kDN, 0, st>>>(yb, lo, up, frm, to, bw, mode_out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
