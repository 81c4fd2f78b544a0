// select.cu — exact order statistics by MSB radix select on 64-bit order-preserving keys.
//
// Replaces (reference side) R stats::median / stats::mad inside abc:::normalise and
// `sort(dist)[nacc]` inside abc::abc / abc::postpr.  Distribution independent:
// six passes (11,11,11,11,11,9 bits).  Each pass = hist kernel -> [optional cross-rank all-reduce of
// the int64 histogram] -> pick kernel.  Up to two ranks per job share one read of the data (the two
// middle order statistics of an even-length column).
#include "common.cuh"

namespace {

constexpr int kBins = ABCDLS_SELECT_BINS;
constexpr int kThreads = 256;

struct SelJob {
    const double* base;
    long long n;
    long long k[2];
    unsigned long long prefix[2];
    long long krem[2];
    double center;
    int use_center;
    int nq;
    int bad;
    int pad;
};

struct SelWs {
    SelJob* jobs;
    unsigned long long* ghist;  // [njobs][2][kBins]
};

__host__ __device__ inline size_t jobs_bytes(int njobs) {
    return ((size_t)njobs * sizeof(SelJob) + 255) / 256 * 256;
}

inline SelWs carve(void* ws, int njobs) {
    SelWs w;
    w.jobs = reinterpret_cast<SelJob*>(ws);
    w.ghist = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + jobs_bytes(njobs));
    return w;
}

__global__ void k_select_setup(SelJob* jobs, int njobs, const double* base, long long n, long long ld,
                               const long long* counts, const double* center, const long long* k0d,
                               const long long* k1d, long long k0h, long long k1h) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    SelJob J;
    J.base = base + (size_t)j * ld;
    J.n = counts ? counts[j] : n;
    J.k[0] = k0d ? k0d[j] : k0h;
    J.k[1] = k1d ? k1d[j] : k1h;
    J.nq = (J.k[1] != J.k[0]) ? 2 : 1;
    J.prefix[0] = J.prefix[1] = 0ull;
    J.krem[0] = J.k[0];
    J.krem[1] = J.k[1];
    J.center = center ? center[j] : 0.0;
    J.use_center = center ? 1 : 0;
    J.bad = 0;
    J.pad = 0;
    jobs[j] = J;
}

__device__ __forceinline__ void hist_add(unsigned* h, unsigned digit, bool active) {
    unsigned am = __ballot_sync(0xffffffffu, active);
    if (active) {
        unsigned peers = __match_any_sync(am, digit);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[digit], (unsigned)__popc(peers));
    }
}

__global__ void __launch_bounds__(kThreads) k_select_hist(const SelJob* jobs, unsigned long long* ghist,
                                                          int pass) {
    __shared__ unsigned sh[2][kBins];
    const SelJob& J = jobs[blockIdx.y];
    const double* __restrict__ base = J.base;
    const long long n = J.n;
    const int nq = J.nq;
    const unsigned long long p0 = J.prefix[0], p1 = J.prefix[1];
    const bool use_center = J.use_center != 0;
    const double c = J.center;
    const bool split = (nq == 2) && (pass > 0) && (p0 != p1);

    for (int b = threadIdx.x; b < 2 * kBins; b += kThreads) (&sh[0][0])[b] = 0u;
    __syncthreads();

    const int shift_digit = pass < 5 ? 64 - 11 * (pass + 1) : 0;
    const unsigned mask = pass < 5 ? 2047u : 511u;
    const int shift_prefix = 64 - 11 * pass;  // only used when pass > 0

    const long long stride = (long long)gridDim.x * kThreads;
    // warp-uniform trip count: iterate over the padded range, predicate on i < n
    const long long n_pad = (n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_pad; i += stride) {
        bool in = i < n;
        double x = in ? ld_stream(base + i) : 0.0;
        double v = use_center ? fabs(__dsub_rn(x, c)) : x;
        unsigned long long key = dkey(v);
        unsigned digit = (unsigned)(key >> shift_digit) & mask;
        bool m0 = in && (pass == 0 || (key >> shift_prefix) == p0);
        hist_add(sh[0], digit, m0);
        if (split) {
            bool m1 = in && ((key >> shift_prefix) == p1);
            hist_add(sh[1], digit, m1);
        }
    }
    __syncthreads();
    unsigned long long* g0 = ghist + ((size_t)blockIdx.y * 2 + 0) * kBins;
    unsigned long long* g1 = g0 + kBins;
    for (int b = threadIdx.x; b < kBins; b += kThreads) {
        unsigned a = sh[0][b];
        if (a) {
            atomicAdd(&g0[b], (unsigned long long)a);
            if (nq == 2 && !split) atomicAdd(&g1[b], (unsigned long long)a);
        }
        if (split) {
            unsigned a1 = sh[1][b];
            if (a1) atomicAdd(&g1[b], (unsigned long long)a1);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_select_pick(SelJob* jobs, unsigned long long* ghist, int pass) {
    const int job = blockIdx.x >> 1, q = blockIdx.x & 1;
    SelJob& J = jobs[job];
    if (q >= J.nq) {
        // keep the unused histogram clean for the next pass
        unsigned long long* g = ghist + ((size_t)job * 2 + q) * kBins;
        for (int b = threadIdx.x; b < kBins; b += kThreads) g[b] = 0ull;
        return;
    }
    unsigned long long* g = ghist + ((size_t)job * 2 + q) * kBins;
    // read the query state before any thread may update it (the writer is a single thread below)
    const long long krem = J.krem[q];
    const unsigned long long prefix_in = J.prefix[q];
    constexpr int per = kBins / kThreads;  // 8
    long long loc[per];
    long long s = 0;
#pragma unroll
    for (int t = 0; t < per; ++t) {
        loc[t] = (long long)g[threadIdx.x * per + t];
        s += loc[t];
    }
    // block exclusive scan (64-bit)
    __shared__ long long wsum[kThreads / 32 + 1];
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
        for (int w2 = 0; w2 < kThreads / 32; ++w2) {
            long long t = wsum[w2];
            wsum[w2] = a;
            a += t;
        }
        wsum[kThreads / 32] = a;
    }
    __syncthreads();
    const long long before = wsum[wid] + incl - s;
    const long long total = wsum[kThreads / 32];
    const int bits = pass < 5 ? 11 : 9;
    if (krem >= before && krem < before + s) {
        long long a = before;
        int b = 0;
        bool found = false;
#pragma unroll
        for (int t = 0; t < per; ++t) {
            if (!found) {
                if (krem >= a + loc[t]) {
                    a += loc[t];
                    b = t + 1;
                } else {
                    found = true;
                }
            }
        }
        J.prefix[q] = (prefix_in << bits) | (unsigned long long)(threadIdx.x * per + b);
        J.krem[q] = krem - a;
    }
    if (threadIdx.x == 0 && (krem < 0 || krem >= total)) J.bad = 1;
    __syncthreads();
    for (int b = threadIdx.x; b < kBins; b += kThreads) g[b] = 0ull;
}

__global__ void k_select_end(const SelJob* jobs, int njobs, double* out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    const SelJob& J = jobs[j];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double a = J.bad ? nan : keyd(J.prefix[0]);
    double b = J.bad ? nan : (J.nq == 2 ? keyd(J.prefix[1]) : a);
    out[2 * j] = a;
    out[2 * j + 1] = b;
}

__global__ void k_median_from_ranks(const double* r2, int n, double scale, double* out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // R median (even n): mean of the two middle values; 0.5*a + 0.5*b == RN((a+b)/2), no overflow
    double m = __dadd_rn(__dmul_rn(0.5, r2[2 * j]), __dmul_rn(0.5, r2[2 * j + 1]));
    out[j] = __dmul_rn(scale, m);
}

}  // namespace

extern "C" {

size_t abcdls_select_workspace_bytes(int njobs) {
    if (njobs < 1) njobs = 1;
    return jobs_bytes(njobs) + (size_t)njobs * 2 * kBins * sizeof(unsigned long long);
}

int abcdls_select_begin(abcdls_handle_t h, void* ws, size_t ws_bytes, int njobs, const double* base,
                        int64_t n, int64_t ld, const int64_t* counts, const double* center,
                        const int64_t* k0_dev, const int64_t* k1_dev, int64_t k0_host, int64_t k1_host,
                        void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && base && njobs >= 1 && n >= 1, "select_begin");
    if (ws_bytes < abcdls_select_workspace_bytes(njobs)) {
        h->err = "select workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SelWs w = carve(ws, njobs);
    ABCDLS_CUDA(h, cudaMemsetAsync(w.ghist, 0, (size_t)njobs * 2 * kBins * sizeof(unsigned long long), st));
    k_select_setup<<<(njobs + 63) / 64, 64, 0, st>>>(w.jobs, njobs, base, n, ld,
                                                     (const long long*)counts, center,
                                                     (const long long*)k0_dev, (const long long*)k1_dev,
                                                     k0_host, k1_host);
    ABCDLS_LAUNCH_CHECK(h);
    // remember the sizing info in the (host-visible) handle-free way: stash n in the first hist call
    return ABCDLS_OK;
}

int abcdls_select_hist(abcdls_handle_t h, void* ws, int njobs, int64_t n_max, int pass, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && pass >= 0 && pass < ABCDLS_SELECT_PASSES && njobs >= 1, "select_hist");
    SelWs w = carve(ws, njobs);
    long long want = (n_max + (long long)kThreads * 16 - 1) / ((long long)kThreads * 16);
    long long cap = (long long)h->sm_count * 8 / njobs;
    if (cap < 1) cap = 1;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    dim3 grid((unsigned)want, (unsigned)njobs);
    k_select_hist<<<grid, kThreads, 0, (cudaStream_t)stream>>>(w.jobs, w.ghist, pass);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_select_pick(abcdls_handle_t h, void* ws, int njobs, int pass, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && pass >= 0 && pass < ABCDLS_SELECT_PASSES && njobs >= 1, "select_pick");
    SelWs w = carve(ws, njobs);
    k_select_pick<<<njobs * 2, kThreads, 0, (cudaStream_t)stream>>>(w.jobs, w.ghist, pass);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_select_hist_buffer(abcdls_handle_t h, void* ws, int njobs, int64_t** buf, size_t* count) {
    ABCDLS_CHECK_ARG(h, h && ws && buf && count && njobs >= 1, "select_hist_buffer");
    SelWs w = carve(ws, njobs);
    *buf = reinterpret_cast<int64_t*>(w.ghist);
    *count = (size_t)njobs * 2 * kBins;
    return ABCDLS_OK;
}

int abcdls_select_end(abcdls_handle_t h, void* ws, int njobs, double* out, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ws && out && njobs >= 1, "select_end");
    SelWs w = carve(ws, njobs);
    k_select_end<<<(njobs + 63) / 64, 64, 0, (cudaStream_t)stream>>>(w.jobs, njobs, out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_median_from_ranks(abcdls_handle_t h, const double* ranks2, int n_cols, double scale,
                             double* out, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ranks2 && out && n_cols >= 1, "median_from_ranks");
    k_median_from_ranks<<<(n_cols + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ranks2, n_cols, scale, out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

size_t abcdls_column_mad_workspace_bytes(int d) {
    return abcdls_select_workspace_bytes(d) + align_up((size_t)2 * d * sizeof(double), 256);
}

// Single-GPU exact path: two full radix selects (median, then median of |x - median|).
int abcdls_column_mad(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, double* med,
                      double* mad, void* ws, size_t ws_bytes, int* status_dev, void* stream) {
    (void)status_dev;
    ABCDLS_CHECK_ARG(h, h && ss && med && mad && ws && n >= 1 && d >= 1 && ld >= n, "column_mad");
    if (ws_bytes < abcdls_column_mad_workspace_bytes(d)) {
        h->err = "column_mad workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    size_t sel_bytes = abcdls_select_workspace_bytes(d);
    double* ranks2 = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + sel_bytes);
    const int64_t k0 = (n - 1) / 2, k1 = n / 2;
    for (int phase = 0; phase < 2; ++phase) {
        int rc = abcdls_select_begin(h, ws, sel_bytes, d, ss, n, ld, nullptr, phase ? med : nullptr,
                                     nullptr, nullptr, k0, k1, stream);
        if (rc) return rc;
        for (int p = 0; p < ABCDLS_SELECT_PASSES; ++p) {
            if ((rc = abcdls_select_hist(h, ws, d, n, p, stream))) return rc;
            if ((rc = abcdls_select_pick(h, ws, d, p, stream))) return rc;
        }
        if ((rc = abcdls_select_end(h, ws, d, ranks2, stream))) return rc;
        if ((rc = abcdls_median_from_ranks(h, ranks2, d, phase ? 1.4826 : 1.0, phase ? mad : med, stream)))
            return rc;
    }
    return ABCDLS_OK;
}

}  // extern "C"
