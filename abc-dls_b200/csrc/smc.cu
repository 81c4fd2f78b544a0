// smc.cu — kernel 4: the table-sized parts of ABC-DLS's SMC range update.
//   oldrange   : per-column min / max over ALL simulated parameter rows (ABC.py:2714-2716, params.min()/max())
//   box filter : rows with every parameter inside [lo_j, hi_j] inclusive, ascending row order
//                (ABC.py:2974-2990, Series.between chained over the columns)
// The P-sized range rules themselves (ABC.py:2881-2946, 3016-3028) are host code (smc.py).
#include "common.cuh"
#include "compact.cuh"

namespace {

constexpr int kMaxP = 64;

__global__ void k_minmax_init(unsigned long long* keys, int P) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < P) {
        keys[j] = ~0ull;      // running min of keys
        keys[P + j] = 0ull;   // running max of keys
    }
}

__global__ void __launch_bounds__(256) k_minmax(const double* __restrict__ a, long long n, long long ld,
                                                unsigned long long* __restrict__ keys, int P) {
    const int j = blockIdx.y;
    const double* __restrict__ col = a + (size_t)j * ld;
    double mn = __longlong_as_double(0x7ff0000000000000ll), mx = -mn;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double x0 = ld_stream(col + i), x1 = ld_stream(col + i + stride), x2 = ld_stream(col + i + 2 * stride),
               x3 = ld_stream(col + i + 3 * stride);
        mn = fmin(fmin(mn, x0), fmin(x1, fmin(x2, x3)));
        mx = fmax(fmax(mx, x0), fmax(x1, fmax(x2, x3)));
    }
    for (; i < n; i += stride) {
        double x = ld_stream(col + i);
        mn = fmin(mn, x);
        mx = fmax(mx, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ double smn[8], smx[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        smn[wid] = mn;
        smx[wid] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            mn = fmin(mn, smn[w]);
            mx = fmax(mx, smx[w]);
        }
        // min / max are order independent => atomics on the order-preserving keys stay deterministic
        atomicMin(&keys[j], dkey(mn));
        atomicMax(&keys[P + j], dkey(mx));
    }
}

__global__ void k_minmax_finish(const unsigned long long* keys, int P, double* out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 2 * P) out[j] = keyd(keys[j]);
}

struct Box {
    double lo[kMaxP], hi[kMaxP];
};

__device__ __forceinline__ bool in_box(const double* __restrict__ a, long long ld, long long i, int P, const Box& B) {
    bool ok = true;
    for (int j = 0; j < P; ++j) {
        double x = a[(size_t)j * ld + i];
        ok = ok && (x >= B.lo[j]) && (x <= B.hi[j]);
    }
    return ok;
}

template <bool WRITE>
__global__ void __launch_bounds__(kTileThreads) k_box_tiles(const double* __restrict__ a, long long n, long long ld,
                                                            int P, const double* __restrict__ lo,
                                                            const double* __restrict__ hi, int* __restrict__ tile_cnt,
                                                            unsigned long long* total,
                                                            const long long* __restrict__ tile_off,
                                                            long long row_offset, long long cap,
                                                            long long* __restrict__ idx_out, long long* n_out,
                                                            int* status) {
    __shared__ Box B;
    __shared__ int sh[33];
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        B.lo[j] = lo[j];
        B.hi[j] = hi[j];
    }
    __syncthreads();
    // thread t of the tile handles rows t, t+256, ... (coalesced); ordering is restored per item slice
    const long long tile0 = (long long)blockIdx.x * kTile;
    bool f[kItems];
    int c = 0;
#pragma unroll
    for (int t = 0; t < kItems; ++t) {
        long long i = tile0 + (long long)t * kTileThreads + threadIdx.x;
        f[t] = (i < n) && in_box(a, ld, i, P, B);
        c += f[t] ? 1 : 0;
    }
    if (!WRITE) {
        int tot;
        block_exclusive_scan(c, sh, &tot);
        if (threadIdx.x == 0) {
            tile_cnt[blockIdx.x] = tot;
            if (tot) atomicAdd(total, (unsigned long long)tot);
        }
    } else {
        long long pos = tile_off[blockIdx.x];
#pragma unroll
        for (int t = 0; t < kItems; ++t) {
            int tot;
            int ex = block_exclusive_scan(f[t] ? 1 : 0, sh, &tot);
            if (f[t] && pos + ex < cap) idx_out[pos + ex] = row_offset + tile0 + (long long)t * kTileThreads + threadIdx.x;
            pos += tot;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            long long tt = (long long)*total;
            if (tt > cap) {
                tt = cap;
                atomicOr(status, ABCDLS_S_CAPACITY);
            }
            *n_out = tt;
        }
    }
}

inline long long ntiles_for(int64_t n) { return (n + kTile - 1) / kTile; }

}  // namespace

extern "C" {

size_t abcdls_smc_workspace_bytes(int64_t n, int P) {
    long long nt = ntiles_for(n);
    return 256 + align_up((size_t)2 * P * 8, 256) + align_up((size_t)nt * sizeof(int), 256) +
           align_up((size_t)nt * sizeof(long long), 256);
}

int abcdls_smc_oldrange(abcdls_handle_t h, const double* param_all, int64_t n, int P, int64_t ld, double* minmax,
                        void* ws, size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && param_all && minmax && ws && n >= 1 && P >= 1 && P <= kMaxP && ld >= n, "smc_oldrange");
    if (ws_bytes < abcdls_smc_workspace_bytes(n, P)) {
        h->err = "smc workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ws) + 256);
    k_minmax_init<<<1, 128, 0, st>>>(keys, P);
    long long bx = (n + 256 * 8 - 1) / (256 * 8);
    long long mx = (long long)h->sm_count * 8 / P;
    if (mx < 1) mx = 1;
    if (bx > mx) bx = mx;
    if (bx < 1) bx = 1;
    k_minmax<<<dim3((unsigned)bx, (unsigned)P), 256, 0, st>>>(param_all, n, ld, keys, P);
    k_minmax_finish<<<1, 128, 0, st>>>(keys, P, minmax);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_smc_box_filter(abcdls_handle_t h, const double* param_all, int64_t n, int P, int64_t ld, const double* lo,
                          const double* hi, int64_t row_offset, int64_t cap, int64_t* idx_out, int64_t* n_out,
                          void* ws, size_t ws_bytes, int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && param_all && lo && hi && idx_out && n_out && ws && status_dev, "smc_box_filter: null");
    ABCDLS_CHECK_ARG(h, n >= 1 && P >= 1 && P <= kMaxP && ld >= n && cap >= 1, "smc_box_filter: shape");
    if (ws_bytes < abcdls_smc_workspace_bytes(n, P)) {
        h->err = "smc workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long nt = ntiles_for(n);
    char* p = reinterpret_cast<char*>(ws);
    unsigned long long* total = reinterpret_cast<unsigned long long*>(p);
    size_t off = 256 + align_up((size_t)2 * P * 8, 256);
    int* tile_cnt = reinterpret_cast<int*>(p + off);
    long long* tile_off = reinterpret_cast<long long*>(p + off + align_up((size_t)nt * sizeof(int), 256));
    ABCDLS_CUDA(h, cudaMemsetAsync(total, 0, 8, st));
    k_box_tiles<false><<<(unsigned)nt, kTileThreads, 0, st>>>(param_all, n, ld, P, lo, hi, tile_cnt, total, nullptr,
                                                             row_offset, cap, nullptr, nullptr, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    k_scan_tiles<<<1, 1024, 0, st>>>(tile_cnt, nt, tile_off);
    ABCDLS_LAUNCH_CHECK(h);
    k_box_tiles<true><<<(unsigned)nt, kTileThreads, 0, st>>>(param_all, n, ld, P, lo, hi, tile_cnt, total, tile_off,
                                                            row_offset, cap, (long long*)idx_out, (long long*)n_out,
                                                            status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
