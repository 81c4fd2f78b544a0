// dist.cu — kernel 1 (MAD scaling fused with the Euclidean distance) and kernel 2 (tolerance cut:
// ordered compaction with R's cumsum cap).
//
// Reference semantics (third-party R `abc` 2.2.1, reached from src/Classes/ABC.py:919,1950,2806;
// SURVEY.md §8a R2/R3):
//     scaled[,j] <- sumstat[,j] / mad_j   (unscaled when mad_j == 0);  t_j likewise
//     sum1 <- 0; for j: sum1 <- sum1 + (scaled[,j] - t_j)^2 ; dist <- sqrt(sum1)
//     wt1 <- dist <= ds ; wt1 <- wt1 & cumsum(wt1) <= nacc
// Bit-exactness: IEEE division by the column constant via a Markstein sequence (proven RN),
// explicit __dsub_rn/__dmul_rn/__dadd_rn (no FMA contraction), j strictly sequential, __dsqrt_rn.
#include "common.cuh"
#include "compact.cuh"

namespace {

constexpr int kMaxD = 64;
constexpr int kDistThreads = 256;

struct ColScale {
    double div[kMaxD];   // divisor actually used (mad, or 1 when mad == 0)
    double rcp[kMaxD];   // RN(1/div)
    double tgt[kMaxD];   // target / div
    double lo[kMaxD];    // |x| inside (lo, hi) -> Markstein path is exact
    double hi[kMaxD];
};

__device__ __forceinline__ void load_scale(ColScale& S, const double* mad, const double* target, int d) {
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        double m = mad[j];
        double dv = (m == 0.0) ? 1.0 : m;
        S.div[j] = dv;
        S.rcp[j] = 1.0 / dv;              // IEEE division, once per block
        S.tgt[j] = (m == 0.0) ? target[j] : target[j] / dv;
        // keep q = x/div, x*rcp and the fma remainders away from subnormal / overflow
        double a = fabs(dv);
        S.lo[j] = a * 1e-280 + 1e-280;
        S.hi[j] = fmin(a * 1e280, 1e280);
    }
}

__device__ __forceinline__ double scaled_value(double x, const ColScale& S, int j) {
    double ax = fabs(x);
    if ((ax > S.lo[j] && ax < S.hi[j]) || x == 0.0) return div_const(x, S.div[j], S.rcp[j]);
    return x / S.div[j];
}

// One thread owns ROWS consecutive rows... kept simple: 2 rows per thread via 16-byte loads when the
// column starts are 16-byte aligned (ld even and base aligned), scalar otherwise.
template <bool VEC2>
__global__ void __launch_bounds__(kDistThreads) k_scale_dist(const double* __restrict__ ss, long long n, int d,
                                                             long long ld, const double* __restrict__ mad,
                                                             const double* __restrict__ target,
                                                             double* __restrict__ dist, int* status) {
    pdl_prologue();
    __shared__ ColScale S;
    load_scale(S, mad, target, d);
    __syncthreads();
    bool nonfinite = false;
    if (VEC2) {
        const long long npair = n >> 1;
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npair; p += stride) {
            double s0 = 0.0, s1 = 0.0;
            const double* col = ss + 2 * p;
#pragma unroll 4
            for (int j = 0; j < d; ++j) {
                double2 x = ld_stream2(col);
                col += ld;
                double tj = S.tgt[j];
                double a = __dsub_rn(scaled_value(x.x, S, j), tj);
                double b = __dsub_rn(scaled_value(x.y, S, j), tj);
                s0 = __dadd_rn(s0, __dmul_rn(a, a));
                s1 = __dadd_rn(s1, __dmul_rn(b, b));
            }
            double2 r;
            r.x = __dsqrt_rn(s0);
            r.y = __dsqrt_rn(s1);
            nonfinite |= !(isfinite(r.x) && isfinite(r.y));
            *reinterpret_cast<double2*>(dist + 2 * p) = r;
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            long long i = n - 1;
            double s0 = 0.0;
            for (int j = 0; j < d; ++j) {
                double a = __dsub_rn(scaled_value(ss[(size_t)j * ld + i], S, j), S.tgt[j]);
                s0 = __dadd_rn(s0, __dmul_rn(a, a));
            }
            double r = __dsqrt_rn(s0);
            nonfinite |= !isfinite(r);
            dist[i] = r;
        }
    } else {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            double s0 = 0.0;
            const double* col = ss + i;
#pragma unroll 4
            for (int j = 0; j < d; ++j) {
                double x = ld_stream(col);
                col += ld;
                double a = __dsub_rn(scaled_value(x, S, j), S.tgt[j]);
                s0 = __dadd_rn(s0, __dmul_rn(a, a));
            }
            double r = __dsqrt_rn(s0);
            nonfinite |= !isfinite(r);
            dist[i] = r;
        }
    }
    if (__any_sync(0xffffffffu, nonfinite) && (threadIdx.x & 31) == 0) atomicOr(status, ABCDLS_S_NONFINITE);
}

__global__ void __launch_bounds__(kTileThreads) k_tile_count_le(const double* __restrict__ dist, long long n,
                                                                const long long* __restrict__ n_dev,
                                                                const double* __restrict__ ds_ptr,
                                                                int* __restrict__ tile_cnt,
                                                                unsigned long long* total) {
    pdl_prologue();
    __shared__ int sh[33];
    const double ds = *ds_ptr;
    if (n_dev && *n_dev < n) n = *n_dev;
    const long long base = (long long)blockIdx.x * kTile + (long long)threadIdx.x * kItems;
    int c = 0;
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        long long i = base + t;
        if (i < n) c += (dist[i] <= ds) ? 1 : 0;
    }
    int tot;
    block_exclusive_scan(c, sh, &tot);
    if (threadIdx.x == 0) {
        tile_cnt[blockIdx.x] = tot;
        if (tot) atomicAdd(total, (unsigned long long)tot);
    }
}

__global__ void __launch_bounds__(kTileThreads) k_tile_write_le(const double* __restrict__ dist, long long n,
                                                                const long long* __restrict__ n_dev,
                                                                const long long* __restrict__ src_idx,
                                                                double* __restrict__ dist_out,
                                                                const double* __restrict__ ds_ptr,
                                                                const long long* __restrict__ tile_off,
                                                                const long long* __restrict__ quota_ptr,
                                                                long long row_offset, long long cap,
                                                                long long* __restrict__ idx_out,
                                                                long long* __restrict__ slot_out,
                                                                const unsigned long long* __restrict__ total,
                                                                long long* __restrict__ n_out, int* status) {
    pdl_prologue();
    __shared__ int sh[33];
    const double ds = *ds_ptr;
    long long quota = *quota_ptr;
    if (quota < 0) quota = 0;
    if (n_dev && *n_dev < n) n = *n_dev;
    const long long base = (long long)blockIdx.x * kTile + (long long)threadIdx.x * kItems;
    bool f[kItems];
    int c = 0;
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        long long i = base + t;
        f[t] = (i < n) && (dist[i] <= ds);
        c += f[t] ? 1 : 0;
    }
    int tot;
    int ex = block_exclusive_scan(c, sh, &tot);
    long long pos = tile_off[blockIdx.x] + ex;
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        if (f[t]) {
            if (pos < quota && pos < cap) {
                idx_out[pos] = src_idx ? src_idx[base + t] : row_offset + base + t;
                if (dist_out) dist_out[pos] = dist[base + t];
                if (slot_out) slot_out[pos] = base + t;
            }
            ++pos;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long tt = (long long)*total;
        long long m = tt < quota ? tt : quota;
        if (m > cap) {
            m = cap;
            atomicOr(status, ABCDLS_S_CAPACITY);
        }
        *n_out = m;
    }
}

}  // namespace

extern "C" {

int abcdls_scale_dist(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* mad,
                      const double* target, double* dist, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && mad && target && dist && status_dev, "scale_dist: null pointer");
    ABCDLS_CHECK_ARG(h, n >= 1 && d >= 1 && d <= kMaxD && ld >= n, "scale_dist: shape (d <= 64)");
    const bool vec = ((reinterpret_cast<uintptr_t>(ss) & 15) == 0) && ((ld & 1) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dist) & 15) == 0);
    long long work = vec ? (n + 1) / 2 : n;
    long long blocks = (work + kDistThreads - 1) / kDistThreads;
    long long maxb = (long long)h->sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    if (vec)
        k_scale_dist<true><<<(unsigned)blocks, kDistThreads, 0, (cudaStream_t)stream>>>(ss, n, d, ld, mad, target,
                                                                                       dist, status_dev);
    else
        k_scale_dist<false><<<(unsigned)blocks, kDistThreads, 0, (cudaStream_t)stream>>>(ss, n, d, ld, mad, target,
                                                                                        dist, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

// workspace: [total u64][pad][tile_cnt int x ntiles][tile_off i64 x ntiles]
static inline long long ntiles_for(int64_t n) { return (n + kTile - 1) / kTile; }

size_t abcdls_accept_workspace_bytes(int64_t n) {
    long long nt = ntiles_for(n);
    return 256 + align_up((size_t)nt * sizeof(int), 256) + align_up((size_t)nt * sizeof(long long), 256);
}

int abcdls_count_le(abcdls_handle_t h, const double* dist, int64_t n, const int64_t* n_dev, const double* ds,
                    int64_t* le_count, void* ws, size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && dist && ds && ws && n >= 1, "count_le");
    if (ws_bytes < abcdls_accept_workspace_bytes(n)) {
        h->err = "accept workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long nt = ntiles_for(n);
    char* p = reinterpret_cast<char*>(ws);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(p);
    int* tile_cnt = reinterpret_cast<int*>(p + 256);
    long long* tile_off = reinterpret_cast<long long*>(p + 256 + align_up((size_t)nt * sizeof(int), 256));
    ABCDLS_CUDA(h, cudaMemsetAsync(total, 0, 8, st));
    launch_pdl(k_tile_count_le, dim3((unsigned)nt), dim3(kTileThreads), 0, st, dist, n, (const long long*)n_dev, ds, tile_cnt, total);
    ABCDLS_LAUNCH_CHECK(h);
    launch_pdl(k_scan_tiles, dim3(1), dim3(1024), 0, st, tile_cnt, nt, tile_off);
    ABCDLS_LAUNCH_CHECK(h);
    if (le_count)
        ABCDLS_CUDA(h, cudaMemcpyAsync(le_count, total, 8, cudaMemcpyDeviceToDevice, st));
    return ABCDLS_OK;
}

int abcdls_accept(abcdls_handle_t h, const double* dist, int64_t n, const int64_t* n_dev, const int64_t* src_idx,
                  const double* ds, const int64_t* quota, int64_t row_offset, int64_t cap, int64_t* idx_out,
                  double* dist_out, int64_t* slot_out, int64_t* n_out, void* ws, size_t ws_bytes, int* status_dev,
                  void* stream) {
    ABCDLS_CHECK_ARG(h, h && dist && ds && quota && idx_out && n_out && ws && status_dev && n >= 1 && cap >= 1,
                     "accept");
    if (ws_bytes < abcdls_accept_workspace_bytes(n)) {
        h->err = "accept workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    long long nt = ntiles_for(n);
    char* p = reinterpret_cast<char*>(ws);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(p);
    long long* tile_off = reinterpret_cast<long long*>(p + 256 + align_up((size_t)nt * sizeof(int), 256));
    launch_pdl(k_tile_write_le, dim3((unsigned)nt), dim3(kTileThreads), 0, (cudaStream_t)stream, 
        dist, n, (const long long*)n_dev, (const long long*)src_idx, dist_out, ds, tile_off,
        (const long long*)quota, row_offset, cap, (long long*)idx_out, (long long*)slot_out, total, (long long*)n_out,
        status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
