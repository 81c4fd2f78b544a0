// loclinear.cu — kernel 3: Epanechnikov-weighted local-linear adjustment on the accepted rows,
// plus the postpr model counts and the column statistics used by summary()/region checks.
//
// Reference semantics: third-party R `abc` 2.2.1 abc(method="loclinear") as reached from
// src/Classes/ABC.py:1950-1953 / 2806-2809:
//     w   <- 1 - (dist[wt1]/ds)^2
//     fit1 <- lsfit(scaled.ss[wt1,], param[wt1,], wt = w);  pred <- [1,t] %*% coef
//     res <- param[wt1,] - [1, scaled.ss[wt1,]] %*% coef;  the_m <- colMeans(res); res <- res - the_m
//     fit2 <- lsfit(scaled.ss[wt1,], log(res^2), wt = w) ; adj <- pred + the_m + pred.sd * res / pred.si
// Design choice: the design matrix is centred on the scaled target, so the intercept IS the
// prediction at the target and X'WX is well conditioned; normal equations in fp64, fixed reduction tree.
#include "common.cuh"

namespace {

constexpr int kMaxD = 64;
constexpr int kMaxP = 64;

struct GatherScale {
    double div[kMaxD], rcp[kMaxD], tgt[kMaxD];
};

// accepted rows -> compact column-major blocks (leading dimension `cap`)
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ ss, long long ld_ss,
                                                const double* __restrict__ param, long long ld_param,
                                                const double* __restrict__ dist_acc, const long long* __restrict__ idx,
                                                const long long* __restrict__ n_acc_ptr, long long row_offset,
                                                long long cap, int d, int P, const double* __restrict__ mad,
                                                const double* __restrict__ target, const double* __restrict__ ds_ptr,
                                                double* __restrict__ xs, double* __restrict__ y,
                                                double* __restrict__ ss_out, double* __restrict__ w) {
    __shared__ GatherScale S;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        double m = mad[j];
        double dv = (m == 0.0) ? 1.0 : m;
        S.div[j] = dv;
        S.tgt[j] = (m == 0.0) ? target[j] : target[j] / dv;
    }
    __syncthreads();
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const double ds = *ds_ptr;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_acc; r += stride) {
        const long long i = idx[r] - row_offset;
        for (int j = 0; j < d; ++j) {
            double x = ss[(size_t)j * ld_ss + i];
            ss_out[(size_t)j * cap + r] = x;
            double sc = (S.div[j] == 1.0) ? x : x / S.div[j];  // IEEE division == R's x/mad
            xs[(size_t)j * cap + r] = __dsub_rn(sc, S.tgt[j]);
        }
        if (param)
            for (int p = 0; p < P; ++p) y[(size_t)p * cap + r] = param[(size_t)p * ld_param + i];
        double q = dist_acc[r] / ds;
        w[r] = __dsub_rn(1.0, __dmul_rn(q, q));
    }
}

// out[p][i] = table[p][rows[i] - row_offset] for the first *n_rows rows (column-major table, possibly pinned host
// memory): the scattered half of a gather, runnable as soon as the row numbers exist.  grid (gx, P)
__global__ void __launch_bounds__(256) k_gather_rows(const double* __restrict__ table, long long ld,
                                                     const long long* __restrict__ rows,
                                                     const long long* __restrict__ n_rows_ptr, long long row_offset,
                                                     long long cap, double* __restrict__ out) {
    long long n = *n_rows_ptr;
    if (n > cap) n = cap;
    const int p = blockIdx.y;
    const double* __restrict__ src = table + (size_t)p * ld - row_offset;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v;
        asm volatile("ld.global.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(src + rows[i]));
        out[(size_t)p * cap + i] = v;
    }
}


// y[p][r] = param[(idx[r] - row_offset) * pitch + p]: ROW-major parameter table (numpy / pandas C order - the P
// parameters of one simulation are contiguous - which is how ABC-DLS's frames lie in memory), possibly pinned host
// memory.  A CTA moves kGprRows accepted rows at a time: consecutive threads read consecutive doubles of a row (one
// 128-byte request per 16 parameters instead of 16 separate sectors of a column-major table), the tile is transposed in
// shared memory, and consecutive threads write consecutive rows of y.
constexpr int kGprRows = 128;
__global__ void __launch_bounds__(256) k_gather_param_rm(const double* param, long long pitch, int P,
                                                         const long long* __restrict__ idx,
                                                         const long long* __restrict__ n_acc_ptr, long long row_offset,
                                                         long long cap, double* __restrict__ y) {
    extern __shared__ double gpr_tile[];   // [P][kGprRows + 1]
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const long long ntile = (n_acc + kGprRows - 1) / kGprRows;
    for (long long tl = blockIdx.x; tl < ntile; tl += gridDim.x) {
        const long long r0 = tl * kGprRows;
        const int nr = (int)min((long long)kGprRows, n_acc - r0);
        const int ne = nr * P;
        for (int e0 = threadIdx.x; e0 < ne; e0 += 256 * 4) {
            double v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {      // four independent loads in flight per thread
                const int e = e0 + k * 256;
                if (e < ne) {
                    const int rr = e / P, pp = e - rr * P;
                    v[k] = param[(size_t)(idx[r0 + rr] - row_offset) * pitch + pp];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int e = e0 + k * 256;
                if (e < ne) {
                    const int rr = e / P, pp = e - rr * P;
                    gpr_tile[pp * (kGprRows + 1) + rr] = v[k];
                }
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < P * kGprRows; e += 256) {
            const int pp = e / kGprRows, rr = e - pp * kGprRows;
            if (rr < nr) y[(size_t)pp * cap + r0 + rr] = gpr_tile[pp * (kGprRows + 1) + rr];
        }
        __syncthreads();
    }
}

// same outputs from the block of candidate rows the one-pass kernel emitted (row-major, `pitch` doubles per slot,
// reached through perm[slot[r]]) and the parameter table (possibly pinned host memory) through idx[r].
// blockIdx.y == 0: one thread per accepted row copies / scales its d statistics and computes its weight;
// blockIdx.y == 1 + p: one thread per accepted row gathers parameter p (P independent scattered loads in flight).
__global__ void __launch_bounds__(256) k_gather_cand(const double* __restrict__ cand_e, long long pitch,
                                                     const long long* __restrict__ slot,
                                                     const long long* __restrict__ perm,
                                                     const double* __restrict__ param, long long ld_param,
                                                     const double* __restrict__ cand_par, long long ld_cpar,
                                                     const double* __restrict__ dist_acc,
                                                     const long long* __restrict__ idx,
                                                     const long long* __restrict__ n_acc_ptr, long long row_offset,
                                                     long long cap, int d, const double* __restrict__ mad,
                                                     const double* __restrict__ target,
                                                     const double* __restrict__ ds_ptr, double* __restrict__ xs,
                                                     double* __restrict__ y, double* __restrict__ ss_out,
                                                     double* __restrict__ w) {
    __shared__ double sdiv[kMaxD], stgt[kMaxD];
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long r0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.y == 0) {
        for (int j = threadIdx.x; j < d; j += blockDim.x) {
            const double m = mad[j];
            const double dv = (m == 0.0) ? 1.0 : m;
            sdiv[j] = dv;
            stgt[j] = (m == 0.0) ? target[j] : target[j] / dv;
        }
        __syncthreads();
        const double ds = *ds_ptr;
        for (long long r = r0; r < n_acc; r += stride) {
            const long long sl = perm ? perm[slot[r]] : slot[r];
            const double* __restrict__ rowp = cand_e + (size_t)sl * pitch;
            for (int j = 0; j < d; ++j) {
                const double x = rowp[j];
                ss_out[(size_t)j * cap + r] = x;
                const double sc = (sdiv[j] == 1.0) ? x : x / sdiv[j];   // IEEE division == R's x/mad
                xs[(size_t)j * cap + r] = __dsub_rn(sc, stgt[j]);
            }
            const double q = dist_acc[r] / ds;
            w[r] = __dsub_rn(1.0, __dmul_rn(q, q));
        }
    } else {
        const int p = blockIdx.y - 1;
        if (cand_par) {   // the candidates' parameters were pre-gathered (abcdls_gather_rows): compact, near-sequential
            const double* __restrict__ cp = cand_par + (size_t)p * ld_cpar;
            for (long long r = r0; r < n_acc; r += stride) y[(size_t)p * cap + r] = cp[slot[r]];
            return;
        }
        const double* __restrict__ src = param + (size_t)p * ld_param - row_offset;
        // scattered 8-byte reads: ask L2 for 64-byte fills instead of the default 128 (halves the DRAM traffic)
        for (long long r = r0; r < n_acc; r += stride) {
            double v;
            asm volatile("ld.global.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(src + idx[r]));
            y[(size_t)p * cap + r] = v;
        }
    }
}

// ---- weighted normal equations ------------------------------------------------------------------
// T_r = [1, xs_r (d), rhs_r (P)] (width W = M + P);  C[a][b] = sum_r w_r * T_r[a] * T_r[b] for a < M, b < W.
// The operands are column-major (k-major), so tiles of 128 rows are staged AS THEY ARE ([column][row], odd pitch:
// conflict-free stores) with 8-byte cp.async copies into a double buffer; nothing is transposed.  Every warp walks
// its own rows of the tile; lane (ablk, bblk) owns the 6 x 4 outputs {a = ablk + nA*i} x {b = bblk + nB*j} in
// registers.  The interleaved index sets make the lanes of one shared load hit consecutive columns (distinct banks)
// while lanes with the same block share one broadcast.  Per row and warp: 10 shared loads, 6 multiplies, 24 FMAs.
// Warps, then CTAs, are summed in a fixed order: deterministic for a given (n_acc, cap).
constexpr int kNormThreads = 256;
constexpr int kNormWarps = kNormThreads / 32;
constexpr int kNormRows = 128;           // rows per staged tile
constexpr int kNormPitch = kNormRows + 1;
constexpr int kBA = 6, kBB = 4;

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kNormThreads) k_normal_partial(const double* __restrict__ xs,
                                                                 const double* __restrict__ rhs,
                                                                 const double* __restrict__ w,
                                                                 const long long* __restrict__ n_acc_ptr,
                                                                 long long cap, int d, int P, int with_gram,
                                                                 const double* __restrict__ the_m,
                                                                 double* __restrict__ res_io,
                                                                 double* __restrict__ partials) {
    // the_m != NULL ("hcorr" fit): rhs = res_io holds the raw residuals; they are centred in place (res -= the_m)
    // and the staged right-hand side becomes log(res^2) - R's fit2 without a log(res^2) round trip through HBM
    extern __shared__ __align__(16) double smn[];
    const int M = d + 1, W = M + P;
    const int b0 = with_gram ? 0 : M;                  // first output column
    const int Wb = W - b0;
    const int nA = (M + kBA - 1) / kBA, nB = (Wb + kBB - 1) / kBB;
    const int npairs = nA * nB;
    const int nslots = (npairs + 31) / 32;
    // <= 16 register blocks (the hcorr refit: no Gram part): the two half-warps take alternate rows instead of leaving
    // 20 lanes idle - half the instructions per tile
    const bool split = npairs <= 16;
    // tile columns: 0 = ones, 1..d = xs, M..W-1 = rhs, W = weights, W + 1 = zeros (padding target)
    const int ncols = W + 2;
    double* buf[2] = {smn, smn + (size_t)ncols * kNormPitch};
    double* red = smn;   // [kNormWarps][32][24], aliases the tile buffers (used only between the tile loops)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const long long ntile = (n_acc + kNormRows - 1) / kNormRows;
    const long long per = (ntile + gridDim.x - 1) / gridDim.x;
    const long long t_begin = (long long)blockIdx.x * per, t_end = min(ntile, t_begin + per);
    const int E = (with_gram ? M * M : 0) + M * P;
    double* out = partials + (size_t)blockIdx.x * E;
    for (int e = threadIdx.x; e < E; e += kNormThreads) out[e] = 0.0;

    auto stage = [&](long long tile, double* T) {
        const long long r0 = tile * kNormRows;
        const int nr = (int)min((long long)kNormRows, n_acc - r0);
        for (int t = threadIdx.x; t < W * kNormRows; t += kNormThreads) {
            const int c = 1 + t / kNormRows, rr = t % kNormRows;       // c = 1 .. W  (W = the weight column)
            double* dst = T + (size_t)c * kNormPitch + rr;
            const double* src = (c < M ? xs + (size_t)(c - 1) * cap : (c < W ? rhs + (size_t)(c - M) * cap : w)) + r0 + rr;
            if (rr < nr) cp_async8(dst, src);
            else *dst = 0.0;
        }
    };

    for (int slot = 0; slot < nslots; ++slot) {
        const int pair = split ? (lane & 15) : slot * 32 + lane;
        const bool on = pair < npairs;
        const int ablk = on ? pair / nB : 0, bblk = on ? pair % nB : 0;
        int ai[kBA], bi[kBB];                                            // tile columns (padding -> the zero column)
#pragma unroll
        for (int i = 0; i < kBA; ++i) {
            const int a = ablk + nA * i;
            ai[i] = (on && a < M) ? a * kNormPitch : (W + 1) * kNormPitch;
        }
#pragma unroll
        for (int j = 0; j < kBB; ++j) {
            const int b = b0 + bblk + nB * j;
            bi[j] = (on && b < W) ? b * kNormPitch : (W + 1) * kNormPitch;
        }
        double acc[kBA][kBB];
#pragma unroll
        for (int i = 0; i < kBA; ++i)
#pragma unroll
            for (int j = 0; j < kBB; ++j) acc[i][j] = 0.0;

        __syncthreads();
        for (int t = threadIdx.x; t < 2 * kNormPitch; t += kNormThreads) {   // constant columns of both buffers
            const int bsel = t / kNormPitch, rr = t % kNormPitch;
            buf[bsel][rr] = 1.0;
            buf[bsel][(size_t)(W + 1) * kNormPitch + rr] = 0.0;
        }
        if (t_begin < t_end) stage(t_begin, buf[0]);
        cp_async_commit();
        for (long long tile = t_begin; tile < t_end; ++tile) {
            const int cur = (int)((tile - t_begin) & 1);
            if (tile + 1 < t_end) stage(tile + 1, buf[cur ^ 1]);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            if (the_m) {
                double* Tm = buf[cur];
                const long long r0 = tile * kNormRows;
                const int nr = (int)min((long long)kNormRows, n_acc - r0);
                for (int t = threadIdx.x; t < P * kNormRows; t += kNormThreads) {
                    const int p = t / kNormRows, rr = t % kNormRows;
                    if (rr < nr) {
                        double* cell = Tm + (size_t)(M + p) * kNormPitch + rr;
                        double v = *cell;
                        if (slot == 0) {               // later slots re-stage residuals that are already centred
                            v -= the_m[p];
                            res_io[(size_t)p * cap + r0 + rr] = v;
                        }
                        *cell = log(v * v);
                    }
                }
                __syncthreads();
            }
            const double* __restrict__ T = buf[cur];
            const double* __restrict__ wt = T + (size_t)W * kNormPitch;
            const int rstart = split ? 2 * wid + (lane >> 4) : wid, rstep = split ? 2 * kNormWarps : kNormWarps;
#pragma unroll 2
            for (int rr = rstart; rr < kNormRows; rr += rstep) {
                const double wv = wt[rr];
                double av[kBA], bv[kBB];
#pragma unroll
                for (int i = 0; i < kBA; ++i) av[i] = T[ai[i] + rr] * wv;
#pragma unroll
                for (int j = 0; j < kBB; ++j) bv[j] = T[bi[j] + rr];
#pragma unroll
                for (int i = 0; i < kBA; ++i)
#pragma unroll
                    for (int j = 0; j < kBB; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();   // everyone is done with buf[cur] before it is refilled
        }
        cp_async_wait<0>();
        __syncthreads();
        // warps -> CTA partial in warp order
        double* mine = red + ((size_t)wid * 32 + lane) * (kBA * kBB);
#pragma unroll
        for (int i = 0; i < kBA; ++i)
#pragma unroll
            for (int j = 0; j < kBB; ++j) mine[i * kBB + j] = acc[i][j];
        __syncthreads();
        for (int t = threadIdx.x; t < 32 * kBA * kBB; t += kNormThreads) {
            const int ln = t / (kBA * kBB), ij = t % (kBA * kBB);
            const int pr = split ? ln : slot * 32 + ln;
            if (pr >= npairs || (split && ln >= 16)) continue;
            const int a = pr / nB + nA * (ij / kBB), b = b0 + pr % nB + nB * (ij % kBB);
            if (a >= M || b >= W) continue;
            double sum = 0.0;
            for (int wq = 0; wq < kNormWarps; ++wq) {
                sum += red[((size_t)wq * 32 + ln) * (kBA * kBB) + ij];
                if (split) sum += red[((size_t)wq * 32 + ln + 16) * (kBA * kBB) + ij];
            }
            const int e = (b < M) ? a * M + b : (with_gram ? M * M : 0) + a * P + (b - M);
            out[e] = sum;
        }
        __syncthreads();
    }
}

// out[e] = sum over the CTA partials, 8 lanes per entry: lane q adds parts q, q + 8, ... in order, then a fixed
// shuffle tree combines the 8 sums => deterministic
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ partials, int nparts, int E,
                                                         double* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = t >> 3, q = t & 7;
    double a = 0.0;
    if (e < E)
        for (int p = q; p < nparts; p += 8) a += partials[(size_t)p * E + e];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o, 8);
    if (e < E && q == 0) out[e] = a;
}

// Cholesky solve of the (M x M) SPD gram against P right-hand sides. beta[a*P + p].
// Column c: thread c finishes the diagonal, then the rows below it are updated in parallel (thread r owns row r).
__global__ void __launch_bounds__(128) k_solve(const double* __restrict__ gram, const double* __restrict__ rhs, int M,
                                               int P, double* __restrict__ beta, int* status) {
    extern __shared__ double sm[];
    double* L = sm;  // M*M
    __shared__ int bad;
    __shared__ double dmax_sh;
    for (int t = threadIdx.x; t < M * M; t += blockDim.x) L[t] = gram[t];
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double dmax = 0.0;
        for (int a = 0; a < M; ++a) dmax = fmax(dmax, fabs(L[a * M + a]));
        dmax_sh = dmax;
    }
    __syncthreads();
    for (int c = 0; c < M; ++c) {
        if (threadIdx.x == 0) {
            double s = L[c * M + c];
            const double orig = s;
            for (int k = 0; k < c; ++k) s -= L[c * M + k] * L[c * M + k];
            if (!(s > 1e-13 * orig) || !(s > 0.0) || !(dmax_sh > 0.0)) {
                bad = 1;
                s = 1.0;
            }
            L[c * M + c] = sqrt(s);
        }
        __syncthreads();
        const double lcc = L[c * M + c];
        for (int r = c + 1 + threadIdx.x; r < M; r += blockDim.x) {
            double t = L[r * M + c];
            for (int k = 0; k < c; ++k) t -= L[r * M + k] * L[c * M + k];
            L[r * M + c] = t / lcc;
        }
        __syncthreads();
    }
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        double z[kMaxD + 1];
        for (int a = 0; a < M; ++a) {
            double t = rhs[a * P + p];
            for (int k = 0; k < a; ++k) t -= L[a * M + k] * z[k];
            z[a] = t / L[a * M + a];
        }
        for (int a = M - 1; a >= 0; --a) {
            double t = z[a];
            for (int k = a + 1; k < M; ++k) t -= L[k * M + a] * z[k];
            z[a] = t / L[a * M + a];
        }
        for (int a = 0; a < M; ++a) beta[a * P + p] = bad ? __longlong_as_double(0x7ff8000000000000ll) : z[a];
    }
    if (threadIdx.x == 0 && bad) atomicOr(status, ABCDLS_S_SINGULAR);
}

// res[r,p] = y[r,p] - (beta[0,p] + sum_j xs[r,j] * beta[j+1,p])
__global__ void __launch_bounds__(256) k_residual(const double* __restrict__ xs, const double* __restrict__ y,
                                                  const double* __restrict__ beta,
                                                  const long long* __restrict__ n_acc_ptr, long long cap, int d,
                                                  int P, double* __restrict__ res) {
    extern __shared__ double sb[];  // (d+1)*P
    for (int t = threadIdx.x; t < (d + 1) * P; t += blockDim.x) sb[t] = beta[t];
    __syncthreads();
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_acc; r += stride) {
        for (int p0 = 0; p0 < P; p0 += 8) {
            double f[8];
            const int pe = min(8, P - p0);
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = q < pe ? sb[p0 + q] : 0.0;
            for (int j = 0; j < d; ++j) {
                double x = xs[(size_t)j * cap + r];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < pe) f[q] = fma(x, sb[(j + 1) * P + p0 + q], f[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < pe) res[(size_t)(p0 + q) * cap + r] = y[(size_t)(p0 + q) * cap + r] - f[q];
        }
    }
}

// Same residuals, 8 parameters per thread (blockIdx.y picks the group), fused with the three column sums the next
// steps need:  stat[0][p] = sum res, stat[1][p] = sum w res, stat[2][p] = sum w res^2   (the_m = stat0 / n;
// sigma2 = (stat2 - 2 m stat1 + m^2 sum w) / sum w).  Block partials are combined by the last block of each group
// in block order => deterministic.
constexpr int kResBlocks = 592;
__global__ void __launch_bounds__(256) k_residual_stats(const double* __restrict__ xs, const double* __restrict__ y,
                                                        const double* __restrict__ beta, const double* __restrict__ w,
                                                        const long long* __restrict__ n_acc_ptr, long long cap, int d,
                                                        int P, double* __restrict__ res, double* __restrict__ part,
                                                        unsigned* __restrict__ done, double* __restrict__ stat) {
    extern __shared__ double sb[];  // (d+1)*8 coefficients of this group | 8 warps x 24 partials
    double* sred = sb + (d + 1) * 8;
    __shared__ bool last;
    const int p0 = blockIdx.y * 8, pe = min(8, P - p0);
    for (int t = threadIdx.x; t < (d + 1) * 8; t += blockDim.x) {
        const int j = t / 8, q = t % 8;
        sb[t] = q < pe ? beta[j * P + p0 + q] : 0.0;
    }
    __syncthreads();
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    double s0[8], s1[8], s2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s0[q] = s1[q] = s2[q] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_acc; r += stride) {
        double f[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = sb[q];
#pragma unroll 8
        for (int j = 0; j < d; ++j) {
            const double x = xs[(size_t)j * cap + r];
#pragma unroll
            for (int q = 0; q < 8; ++q) f[q] = fma(x, sb[(j + 1) * 8 + q], f[q]);
        }
        const double wr = w[r];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < pe) {
                const double v = y[(size_t)(p0 + q) * cap + r] - f[q];
                res[(size_t)(p0 + q) * cap + r] = v;
                s0[q] += v;
                s1[q] = fma(wr, v, s1[q]);
                s2[q] = fma(wr * v, v, s2[q]);
            }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        s0[q] = warp_sum(s0[q]);
        s1[q] = warp_sum(s1[q]);
        s2[q] = warp_sum(s2[q]);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            sred[wid * 24 + q] = s0[q];
            sred[wid * 24 + 8 + q] = s1[q];
            sred[wid * 24 + 16 + q] = s2[q];
        }
    }
    __syncthreads();
    // part[(y * gridDim.x + x) * 24 + k*8 + q]
    if (threadIdx.x < 24) {
        double a = 0.0;
        for (int wq = 0; wq < 8; ++wq) a += sred[wq * 24 + threadIdx.x];
        part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 24 + threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&done[blockIdx.y], 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) {
        // 10 threads per output sum the block partials b = g, g + 10, ... in order; the 10 sums are added in order
        __threadfence();
        const int o = threadIdx.x % 24, g = threadIdx.x / 24;   // g = 0..9 (threads 240..255 idle)
        double a = 0.0;
        if (g < 10)
            for (unsigned b = g; b < gridDim.x; b += 10)
                a += reinterpret_cast<const volatile double*>(part)[((size_t)blockIdx.y * gridDim.x + b) * 24 + o];
        __syncthreads();
        if (g < 10) sred[g * 24 + o] = a;
        __syncthreads();
        if (threadIdx.x < 24) {
            double t = 0.0;
            for (int gg = 0; gg < 10; ++gg) t += sred[gg * 24 + threadIdx.x];
            const int k = threadIdx.x / 8, q = threadIdx.x % 8;
            if (q < pe) stat[(size_t)k * P + p0 + q] = t;
        }
        if (threadIdx.x == 0) done[blockIdx.y] = 0u;   // ready for the next call on this workspace
    }
}

__global__ void __launch_bounds__(256) k_recentre_log(double* __restrict__ res, const double* __restrict__ the_m,
                                                      const long long* __restrict__ n_acc_ptr, long long cap, int P,
                                                      double* __restrict__ logres2) {
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const int p = blockIdx.y;
    const double m = the_m[p];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_acc; r += stride) {
        double v = res[(size_t)p * cap + r] - m;
        res[(size_t)p * cap + r] = v;
        if (logres2) logres2[(size_t)p * cap + r] = log(v * v);
    }
}

__global__ void __launch_bounds__(256) k_adjust(const double* __restrict__ xs, double* __restrict__ res,
                                                const double* __restrict__ beta, const double* __restrict__ the_m,
                                                const double* __restrict__ gamma,
                                                const long long* __restrict__ n_acc_ptr, long long cap, int d, int P,
                                                int hcorr, double* __restrict__ adj) {
    extern __shared__ double sg[];  // (d+1)*P gamma
    if (hcorr)
        for (int t = threadIdx.x; t < (d + 1) * P; t += blockDim.x) sg[t] = gamma[t];
    __syncthreads();
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_acc; r += stride) {
        double xr[16];
        const bool small = hcorr && d <= 16;     // the row's statistics stay in registers for all P parameters
        if (small) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xr[j] = j < d ? xs[(size_t)j * cap + r] : 0.0;
        }
        for (int p = 0; p < P; ++p) {
            double rv = res[(size_t)p * cap + r];
            if (hcorr) {
                double g0 = sg[p];
                double lin = g0;
                if (small) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < d) lin = fma(xr[j], sg[(j + 1) * P + p], lin);
                } else {
                    for (int j = 0; j < d; ++j) lin = fma(xs[(size_t)j * cap + r], sg[(j + 1) * P + p], lin);
                }
                // pred.sd * res / pred.si = res * sqrt(exp(g0)) / sqrt(exp(lin)) = res * exp((g0 - lin) / 2): one exp
                rv = rv * exp(0.5 * (g0 - lin));
                res[(size_t)p * cap + r] = rv;
            }
            adj[(size_t)p * cap + r] = (beta[p] + the_m[p]) + rv;
        }
    }
}

// per-column statistics over the first n_acc rows: out[c*6 + {0:min,1:max,2:sum,3:sum w*x,4:sum w*x*x,5:sum w}]
// grid (kStatChunks, ncols): every block reduces a fixed contiguous chunk, the last block to finish a column
// combines the chunk partials in chunk order => deterministic for a given (n_acc, cap).
constexpr int kStatChunks = 64;

__device__ __forceinline__ void stat_reduce6(double v[6], double (*sh)[32]) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v[0] = fmin(v[0], __shfl_xor_sync(0xffffffffu, v[0], o));
        v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], o));
#pragma unroll
        for (int q = 2; q < 6; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    }
    __syncthreads();
    if (lane == 0)
        for (int q = 0; q < 6; ++q) sh[q][wid] = v[q];
    __syncthreads();
    if (wid == 0) {
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        v[0] = lane < nw ? sh[0][lane] : inf;
        v[1] = lane < nw ? sh[1][lane] : -inf;
        for (int q = 2; q < 6; ++q) v[q] = lane < nw ? sh[q][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v[0] = fmin(v[0], __shfl_xor_sync(0xffffffffu, v[0], o));
            v[1] = fmax(v[1], __shfl_xor_sync(0xffffffffu, v[1], o));
#pragma unroll
            for (int q = 2; q < 6; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        }
    }
}

__global__ void __launch_bounds__(256) k_col_stats(const double* __restrict__ v, const double* __restrict__ w,
                                                   const long long* __restrict__ n_acc_ptr, long long cap,
                                                   double* __restrict__ part, unsigned* __restrict__ done,
                                                   double* __restrict__ out) {
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const int c = blockIdx.y;
    const double* col = v + (size_t)c * cap;
    const long long per = (n_acc + kStatChunks - 1) / kStatChunks;
    const long long lo = (long long)blockIdx.x * per, hi = min(n_acc, lo + per);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double a[6] = {inf, -inf, 0, 0, 0, 0};
    for (long long r = lo + threadIdx.x; r < hi; r += blockDim.x) {
        const double x = col[r];
        const double ww = w ? w[r] : 1.0;
        a[0] = fmin(a[0], x);
        a[1] = fmax(a[1], x);
        a[2] += x;
        a[3] = fma(ww, x, a[3]);
        a[4] = fma(ww * x, x, a[4]);
        a[5] += ww;
    }
    __shared__ double sh[6][32];
    __shared__ bool last;
    stat_reduce6(a, sh);
    if (threadIdx.x == 0) {
        double* p6 = part + ((size_t)c * kStatChunks + blockIdx.x) * 6;
        for (int q = 0; q < 6; ++q) p6[q] = a[q];
        __threadfence();
        last = (atomicAdd(&done[c], 1u) == (unsigned)kStatChunks - 1);
    }
    __syncthreads();
    if (last) {
        // thread k < 64 holds chunk k; fixed shuffle tree inside each of the two warps, then warp 0 + warp 1
        __threadfence();
        double r6[6] = {inf, -inf, 0, 0, 0, 0};
        if (threadIdx.x < kStatChunks) {
            const volatile double* p6 = part + ((size_t)c * kStatChunks + threadIdx.x) * 6;
            for (int q = 0; q < 6; ++q) r6[q] = p6[q];
        }
        if (threadIdx.x < 64) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                r6[0] = fmin(r6[0], __shfl_down_sync(0xffffffffu, r6[0], o));
                r6[1] = fmax(r6[1], __shfl_down_sync(0xffffffffu, r6[1], o));
#pragma unroll
                for (int q = 2; q < 6; ++q) r6[q] += __shfl_down_sync(0xffffffffu, r6[q], o);
            }
            if ((threadIdx.x & 31) == 0)
                for (int q = 0; q < 6; ++q) sh[q][threadIdx.x >> 5] = r6[q];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            out[(size_t)c * 6 + 0] = fmin(sh[0][0], sh[0][1]);
            out[(size_t)c * 6 + 1] = fmax(sh[1][0], sh[1][1]);
            for (int q = 2; q < 6; ++q) out[(size_t)c * 6 + q] = sh[q][0] + sh[q][1];
            done[c] = 0u;   // ready for the next call on this workspace
        }
    }
}

__global__ void __launch_bounds__(256) k_postpr_counts(const int* __restrict__ model_id,
                                                       const long long* __restrict__ idx,
                                                       const long long* __restrict__ n_acc_ptr, long long row_offset,
                                                       long long cap, int K, unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int shc[];
    for (int t = threadIdx.x; t < K; t += blockDim.x) shc[t] = 0u;
    __syncthreads();
    long long n_acc = *n_acc_ptr;
    if (n_acc > cap) n_acc = cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_acc; r += stride) {
        int m = model_id[idx[r] - row_offset];
        if (m >= 0 && m < K) atomicAdd(&shc[m], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < K; t += blockDim.x)
        if (shc[t]) atomicAdd(&counts[t], (unsigned long long)shc[t]);
}

inline unsigned grid_for(abcdls_handle_t h, long long n, int threads, int per_sm) {
    long long b = (n + threads - 1) / threads;
    long long mx = (long long)h->sm_count * per_sm;
    if (b > mx) b = mx;
    if (b < 1) b = 1;
    return (unsigned)b;
}

inline int normal_grid(abcdls_handle_t h, long long cap) {
    long long b = (cap + kNormRows - 1) / kNormRows;
    long long mx = (long long)h->sm_count * 2;
    if (b > mx) b = mx;
    if (b < 1) b = 1;
    return (int)b;
}


// ---- the P-sized glue between the adjustment kernels (was ~35 one-block torch kernels per call) ------------------------
// rs: [P][6] from k_col_stats(res, w).  sums = [sum res (P) | sum w res (P) | sum w res^2 (P) | n_acc]: the message of
// the cross-rank sum (SURVEY 8e "residual mean the_m").
__global__ void k_resid_moments(const double* __restrict__ rs, const long long* __restrict__ n_acc, int P,
                                double* __restrict__ sums) {
    const int p = threadIdx.x;
    if (p < P) {
        sums[p] = rs[p * 6 + 2];
        sums[P + p] = rs[p * 6 + 3];
        sums[2 * P + p] = rs[p * 6 + 4];
    }
    if (p == 0) sums[3 * P] = (double)*n_acc;
}

// the_m = colMeans(residuals) (UNWEIGHTED, R), sigma2 = sum w (res - the_m)^2 / sum w expanded around the raw moments,
// logsig = sum_p log sigma2_p (aic / bic).  gram[0] = sum w.  No FMA contraction: same roundings as the host formula.
__global__ void k_resid_finish(const double* __restrict__ sums, const double* __restrict__ gram, int P,
                               double* __restrict__ the_m, double* __restrict__ sig, double* __restrict__ logsig) {
    __shared__ double lg[kMaxP];
    const int p = threadIdx.x;
    if (p < P) {
        const double sw = gram[0];
        const double m = __ddiv_rn(sums[p], sums[3 * P]);
        const double a = __dsub_rn(sums[2 * P + p], __dmul_rn(__dmul_rn(2.0, m), sums[P + p]));
        const double s2 = __ddiv_rn(__dadd_rn(a, __dmul_rn(__dmul_rn(m, m), sw)), sw);
        the_m[p] = m;
        sig[p] = s2;
        lg[p] = log(s2);
    }
    __syncthreads();
    if (p == 0) {
        double t = 0.0;
        for (int q = 0; q < P; ++q) t += lg[q];
        *logsig = t;
    }
}

// Multi-rank: everything that is a min, a max (negated) or an OR (negated bits) in ONE vector for one min-exchange,
// the sums of the summary in a second one.  pack = [region min (d) | -region max (d) | posterior min (P) | -posterior
// max (P) | -status bits (8)], sm = stats[:, 2:6] (P x 4).
__global__ void k_final_pack(const int* __restrict__ status, const double* __restrict__ reg,
                             const double* __restrict__ stats, int d, int P, double* __restrict__ pack,
                             double* __restrict__ sm) {
    const int t = threadIdx.x;
    if (t < d) {
        pack[t] = reg[t * 6];
        pack[d + t] = -reg[t * 6 + 1];
    }
    if (t < P) {
        pack[2 * d + t] = stats[t * 6];
        pack[2 * d + P + t] = -stats[t * 6 + 1];
#pragma unroll
        for (int k = 0; k < 4; ++k) sm[t * 4 + k] = stats[t * 6 + 2 + k];
    }
    if (t < 8) pack[2 * d + 2 * P + t] = -(double)((*status >> t) & 1);
}

// All small results of a call in ONE buffer (one copy-out, one blocking read of its head):
// small = [host (8): status, n_acc, ds, region constant?, every MAD zero?, logsig, 0, 0 | mad (d) | stats (P x 6) |
//          sigma2 (P) | pred (P) | coef (M x P) | coef_h (M x P)]      (the last four only for loclinear)
// coef / coef_h: the fits were made on a design centred on the scaled target; R's convention is [1, scaled ss], so the
// intercept moves by -sum_j t_j b_j.  pred = intercept + the_m.
__global__ void k_final_small(const int* __restrict__ status, const long long* __restrict__ n_acc,
                              const double* __restrict__ ds, const double* __restrict__ reg,
                              const double* __restrict__ stats, const double* __restrict__ mad,
                              const double* __restrict__ target, const double* __restrict__ logsig,
                              const double* __restrict__ beta, const double* __restrict__ gamma,
                              const double* __restrict__ the_m, const double* __restrict__ sig,
                              const double* __restrict__ pack, const double* __restrict__ sm, int d, int P,
                              double* __restrict__ small) {
    __shared__ double tsc[kMaxD];
    __shared__ int flags[2];
    const int t = threadIdx.x, M = d + 1;
    double* host = small;
    double* o_mad = small + 8;
    double* o_stats = o_mad + d;
    double* o_sig = o_stats + 6 * P;
    double* o_pred = o_sig + P;
    double* o_coef = o_pred + P;
    double* o_coefh = o_coef + M * P;
    if (t < 2) flags[t] = 1;
    __syncthreads();
    if (t < d) {
        const double m = mad[t];
        o_mad[t] = m;
        tsc[t] = (m == 0.0) ? target[t] : __ddiv_rn(target[t], m);
        const double lo = pack ? pack[t] : reg[t * 6], hi = pack ? -pack[d + t] : reg[t * 6 + 1];
        if (!(lo == hi)) atomicAnd(&flags[0], 0);
        if (m != 0.0) atomicAnd(&flags[1], 0);
    }
    if (t < P) {
        o_stats[t * 6] = pack ? pack[2 * d + t] : stats[t * 6];
        o_stats[t * 6 + 1] = pack ? -pack[2 * d + P + t] : stats[t * 6 + 1];
#pragma unroll
        for (int k = 0; k < 4; ++k) o_stats[t * 6 + 2 + k] = sm ? sm[t * 4 + k] : stats[t * 6 + 2 + k];
    }
    __syncthreads();
    if (beta && t < P) {
        o_sig[t] = sig[t];
        o_pred[t] = beta[t] + the_m[t];
        double acc = 0.0, acch = 0.0;
        for (int j = 0; j < d; ++j) {
            const double b = beta[(j + 1) * P + t];
            o_coef[(j + 1) * P + t] = b;
            acc = __dadd_rn(acc, __dmul_rn(tsc[j], b));
            if (gamma) {
                const double g = gamma[(j + 1) * P + t];
                o_coefh[(j + 1) * P + t] = g;
                acch = __dadd_rn(acch, __dmul_rn(tsc[j], g));
            }
        }
        o_coef[t] = beta[t] - acc;
        if (gamma) o_coefh[t] = gamma[t] - acch;
    }
    if (t == 0) {
        int st = *status;
        if (pack) {
            st = 0;
            for (int b = 0; b < 8; ++b)
                if (pack[2 * d + 2 * P + b] != 0.0) st |= 1 << b;
        }
        host[0] = (double)st;
        host[1] = (double)*n_acc;
        host[2] = *ds;
        host[3] = (double)flags[0];
        host[4] = (double)flags[1];
        host[5] = logsig ? *logsig : 0.0;
        host[6] = host[7] = 0.0;
    }
}

}  // namespace

extern "C" {

int abcdls_gather(abcdls_handle_t h, const double* ss, int64_t ld_ss, const double* param, int64_t ld_param,
                  const double* dist_acc, const int64_t* idx, const int64_t* n_acc, int64_t row_offset, int64_t cap,
                  int d, int P, const double* mad, const double* target, const double* ds, double* xs, double* y,
                  double* ss_out, double* w, void* stream) {
    ABCDLS_CHECK_ARG(h, h && ss && dist_acc && idx && n_acc && mad && target && ds && xs && y && ss_out && w,
                     "gather: null pointer");   // param == NULL: the parameters come from abcdls_gather_param_rows
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP && cap >= 1, "gather: shape");
    k_gather<<<grid_for(h, cap, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        ss, ld_ss, param, ld_param, dist_acc, (const long long*)idx, (const long long*)n_acc, row_offset, cap, d, P,
        mad, target, ds, xs, y, ss_out, w);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_gather_rows(abcdls_handle_t h, const double* table, int64_t ld, int P, const int64_t* rows,
                       const int64_t* n_rows, int64_t row_offset, int64_t cap, double* out, void* stream) {
    ABCDLS_CHECK_ARG(h, h && table && rows && n_rows && out && P >= 1 && P <= kMaxP && cap >= 1, "gather_rows");
    unsigned gx = grid_for(h, cap, 256, 16);
    if (gx > 2048) gx = 2048;
    k_gather_rows<<<dim3(gx, (unsigned)P), 256, 0, (cudaStream_t)stream>>>(table, ld, (const long long*)rows,
                                                                         (const long long*)n_rows, row_offset, cap, out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_gather_param_rows(abcdls_handle_t h, const double* param, int64_t row_pitch, int P, const int64_t* idx,
                             const int64_t* n_acc, int64_t row_offset, int64_t cap, double* y, void* stream) {
    ABCDLS_CHECK_ARG(h, h && param && idx && n_acc && y, "gather_param_rows: null pointer");
    ABCDLS_CHECK_ARG(h, P >= 1 && P <= kMaxP && row_pitch >= P && cap >= 1, "gather_param_rows: shape");
    long long ntile = (cap + kGprRows - 1) / kGprRows;
    unsigned gx = (unsigned)std::min<long long>(ntile, (long long)h->sm_count * 8);
    const size_t smem = (size_t)P * (kGprRows + 1) * sizeof(double);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        ABCDLS_CUDA(h, cudaFuncSetAttribute(k_gather_param_rm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    k_gather_param_rm<<<gx, 256, smem, (cudaStream_t)stream>>>(param, row_pitch, P, (const long long*)idx,
                                                              (const long long*)n_acc, row_offset, cap, y);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_gather_cand(abcdls_handle_t h, const double* cand_ss, int64_t ld_cand, const int64_t* slot,
                       const int64_t* perm, const double* param, int64_t ld_param, const double* cand_par,
                       int64_t ld_cpar, const double* dist_acc, const int64_t* idx,
                       const int64_t* n_acc, int64_t row_offset, int64_t cap, int d, int P, const double* mad,
                       const double* target, const double* ds, double* xs, double* y, double* ss_out, double* w,
                       void* stream) {
    ABCDLS_CHECK_ARG(h, h && cand_ss && slot && dist_acc && idx && n_acc && mad && target && ds && xs && y &&
                            ss_out && w,
                     "gather_cand: null pointer");   // param == NULL: the parameters come from abcdls_gather_param_rows
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP && cap >= 1 && ld_cand >= 1, "gather_cand: shape");
    unsigned gx = grid_for(h, cap, 256, 16);
    if (gx > 2048) gx = 2048;
    k_gather_cand<<<dim3(gx, (unsigned)(1 + ((param || cand_par) ? P : 0))), 256, 0, (cudaStream_t)stream>>>(
        cand_ss, ld_cand, (const long long*)slot, (const long long*)perm, param, ld_param, cand_par, ld_cpar, dist_acc,
        (const long long*)idx,
        (const long long*)n_acc, row_offset, cap, d, mad, target, ds, xs, y, ss_out, w);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

size_t abcdls_normal_workspace_bytes(int d, int P) {
    int M = d + 1;
    size_t E = (size_t)M * M + (size_t)M * P;
    return align_up((size_t)148 * 2 * 2 * E * sizeof(double), 256);  // up to 592 partial blocks
}

static int normal_impl(abcdls_handle_t h, const double* xs, const double* rhs, const double* w, const int64_t* n_acc,
                       int64_t cap, int d, int P, int with_gram, const double* the_m, double* partial, void* ws,
                       size_t ws_bytes, void* stream);

int abcdls_normal(abcdls_handle_t h, const double* xs, const double* rhs, const double* w, const int64_t* n_acc,
                  int64_t cap, int d, int P, int with_gram, double* partial, void* ws, size_t ws_bytes,
                  void* stream) {
    return normal_impl(h, xs, rhs, w, n_acc, cap, d, P, with_gram, nullptr, partial, ws, ws_bytes, stream);
}

int abcdls_normal_logres(abcdls_handle_t h, const double* xs, double* res, const double* the_m, const double* w,
                         const int64_t* n_acc, int64_t cap, int d, int P, double* partial, void* ws, size_t ws_bytes,
                         void* stream) {
    ABCDLS_CHECK_ARG(h, the_m, "normal_logres: null pointer");
    return normal_impl(h, xs, res, w, n_acc, cap, d, P, 0, the_m, partial, ws, ws_bytes, stream);
}

static int normal_impl(abcdls_handle_t h, const double* xs, const double* rhs, const double* w, const int64_t* n_acc,
                       int64_t cap, int d, int P, int with_gram, const double* the_m, double* partial, void* ws,
                       size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && xs && rhs && w && n_acc && partial && ws, "normal: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP && cap >= 1, "normal: shape");
    const int M = d + 1;
    const int E = (with_gram ? M * M : 0) + M * P;
    const int grid = normal_grid(h, cap);
    if (ws_bytes < (size_t)grid * E * sizeof(double)) {
        h->err = "normal workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    size_t smem = (size_t)2 * (M + P + 2) * kNormPitch;
    if (smem < (size_t)kNormWarps * 32 * kBA * kBB) smem = (size_t)kNormWarps * 32 * kBA * kBB;
    smem *= sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    static size_t smem_set = 48 * 1024;   // raise the opt-in limit only when a call needs more than any before
    if (smem > smem_set) {
        ABCDLS_CUDA(h, cudaFuncSetAttribute(k_normal_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    k_normal_partial<<<grid, kNormThreads, smem, st>>>(xs, rhs, w, (const long long*)n_acc, cap, d, P, with_gram, the_m,
                                                       the_m ? const_cast<double*>(rhs) : nullptr,
                                                       reinterpret_cast<double*>(ws));
    ABCDLS_LAUNCH_CHECK(h);
    k_reduce_partials<<<(E * 8 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const double*>(ws), grid, E, partial);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_solve(abcdls_handle_t h, const double* gram, const double* rhs, int d, int P, double* beta,
                 int* status_dev, void* stream) {
    ABCDLS_CHECK_ARG(h, h && gram && rhs && beta && status_dev, "solve: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP, "solve: shape");
    const int M = d + 1;
    k_solve<<<1, 128, (size_t)M * M * sizeof(double), (cudaStream_t)stream>>>(gram, rhs, M, P, beta, status_dev);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_residual(abcdls_handle_t h, const double* xs, const double* y, const double* beta, const int64_t* n_acc,
                    int64_t cap, int d, int P, double* res, void* stream) {
    ABCDLS_CHECK_ARG(h, h && xs && y && beta && n_acc && res, "residual: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP && cap >= 1, "residual: shape");
    k_residual<<<grid_for(h, cap, 256, 8), 256, (size_t)(d + 1) * P * sizeof(double), (cudaStream_t)stream>>>(
        xs, y, beta, (const long long*)n_acc, cap, d, P, res);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

size_t abcdls_residual_workspace_bytes(int P) {
    const int groups = (P + 7) / 8;
    return align_up((size_t)groups * kResBlocks * 24 * sizeof(double), 256) + align_up((size_t)groups * 4, 256);
}

// ws must be zero-initialised once by the caller (the kernel leaves its counters zeroed again).
// stat: [3][P] = {sum res, sum w res, sum w res^2} over the first n_acc rows
int abcdls_residual_stats(abcdls_handle_t h, const double* xs, const double* y, const double* beta, const double* w,
                          const int64_t* n_acc, int64_t cap, int d, int P, double* res, double* stat, void* ws,
                          size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && xs && y && beta && w && n_acc && res && stat && ws, "residual_stats: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP && cap >= 1, "residual_stats: shape");
    if (ws_bytes < abcdls_residual_workspace_bytes(P)) {
        h->err = "residual workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    const int groups = (P + 7) / 8;
    double* part = reinterpret_cast<double*>(ws);
    unsigned* done = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) +
                                                 align_up((size_t)groups * kResBlocks * 24 * sizeof(double), 256));
    unsigned gx = grid_for(h, cap, 256, 2);   // fat blocks: the 24-value block reduction is a fixed cost per block
    if (gx > (unsigned)kResBlocks) gx = kResBlocks;
    const size_t smem = (size_t)((d + 1) * 8 + 10 * 24) * sizeof(double);
    k_residual_stats<<<dim3(gx, groups), 256, smem, (cudaStream_t)stream>>>(xs, y, beta, w, (const long long*)n_acc, cap,
                                                                          d, P, res, part, done, stat);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_recentre_log(abcdls_handle_t h, double* res, const double* the_m, const int64_t* n_acc, int64_t cap,
                        int P, double* logres2, void* stream) {
    ABCDLS_CHECK_ARG(h, h && res && the_m && n_acc, "recentre_log: null pointer");
    ABCDLS_CHECK_ARG(h, P >= 1 && P <= kMaxP && cap >= 1, "recentre_log: shape");
    dim3 grid(grid_for(h, cap, 256, 4), (unsigned)P);
    k_recentre_log<<<grid, 256, 0, (cudaStream_t)stream>>>(res, the_m, (const long long*)n_acc, cap, P, logres2);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_adjust(abcdls_handle_t h, const double* xs, double* res, const double* beta, const double* the_m,
                  const double* gamma, const int64_t* n_acc, int64_t cap, int d, int P, int hcorr, double* adj,
                  void* stream) {
    ABCDLS_CHECK_ARG(h, h && xs && res && beta && the_m && n_acc && adj && (!hcorr || gamma), "adjust: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP && cap >= 1, "adjust: shape");
    k_adjust<<<grid_for(h, cap, 256, 8), 256, (size_t)(d + 1) * P * sizeof(double), (cudaStream_t)stream>>>(
        xs, res, beta, the_m, gamma, (const long long*)n_acc, cap, d, P, hcorr, adj);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

size_t abcdls_col_stats_workspace_bytes(int ncols) {
    return align_up((size_t)ncols * kStatChunks * 6 * sizeof(double), 256) + align_up((size_t)ncols * 4, 256);
}

// ws must be zero-initialised once by the caller (the kernel leaves it zeroed again)
int abcdls_col_stats(abcdls_handle_t h, const double* values, const double* w, const int64_t* n_acc, int64_t cap,
                     int ncols, double* out, void* ws, size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && values && n_acc && out && ws && ncols >= 1 && cap >= 1, "col_stats");
    if (ws_bytes < abcdls_col_stats_workspace_bytes(ncols)) {
        h->err = "col_stats workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    double* part = reinterpret_cast<double*>(ws);
    unsigned* done = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws) +
                                                 align_up((size_t)ncols * kStatChunks * 6 * sizeof(double), 256));
    k_col_stats<<<dim3(kStatChunks, ncols), 256, 0, (cudaStream_t)stream>>>(values, w, (const long long*)n_acc, cap,
                                                                            part, done, out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}


int abcdls_resid_moments(abcdls_handle_t h, const double* rs, const int64_t* n_acc, int P, double* sums, void* stream) {
    ABCDLS_CHECK_ARG(h, h && rs && n_acc && sums && P >= 1 && P <= kMaxP, "resid_moments");
    k_resid_moments<<<1, kMaxP, 0, (cudaStream_t)stream>>>(rs, (const long long*)n_acc, P, sums);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_resid_finish(abcdls_handle_t h, const double* sums, const double* gram, int P, double* the_m, double* sig,
                        double* logsig, void* stream) {
    ABCDLS_CHECK_ARG(h, h && sums && gram && the_m && sig && logsig && P >= 1 && P <= kMaxP, "resid_finish");
    k_resid_finish<<<1, kMaxP, 0, (cudaStream_t)stream>>>(sums, gram, P, the_m, sig, logsig);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

size_t abcdls_final_small_doubles(int d, int P, int loclinear) {
    return (size_t)8 + d + 6 * (size_t)P + (loclinear ? 2 * (size_t)P + 2 * (size_t)(d + 1) * P : 0);
}

int abcdls_final_pack(abcdls_handle_t h, const int* status_dev, const double* reg, const double* stats, int d, int P,
                      double* pack, double* sm, void* stream) {
    ABCDLS_CHECK_ARG(h, h && status_dev && reg && stats && pack && sm, "final_pack: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP, "final_pack: shape");
    k_final_pack<<<1, 64, 0, (cudaStream_t)stream>>>(status_dev, reg, stats, d, P, pack, sm);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_final_small(abcdls_handle_t h, const int* status_dev, const int64_t* n_acc, const double* ds,
                       const double* reg, const double* stats, const double* mad, const double* target,
                       const double* logsig, const double* beta, const double* gamma, const double* the_m,
                       const double* sig, const double* pack, const double* sm, int d, int P, double* small_out,
                       void* stream) {
    ABCDLS_CHECK_ARG(h, h && status_dev && n_acc && ds && reg && stats && mad && target && small_out,
                     "final_small: null pointer");
    ABCDLS_CHECK_ARG(h, d >= 1 && d <= kMaxD && P >= 1 && P <= kMaxP, "final_small: shape");
    ABCDLS_CHECK_ARG(h, !beta || (the_m && sig), "final_small: beta without the_m / sigma2");
    ABCDLS_CHECK_ARG(h, (pack == nullptr) == (sm == nullptr), "final_small: pack and sm come together");
    k_final_small<<<1, 64, 0, (cudaStream_t)stream>>>(status_dev, (const long long*)n_acc, ds, reg, stats, mad, target,
                                                       logsig, beta, gamma, the_m, sig, pack, sm, d, P, small_out);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

int abcdls_postpr_counts(abcdls_handle_t h, const int32_t* model_id, const int64_t* idx, const int64_t* n_acc,
                         int64_t row_offset, int64_t cap, int K, int64_t* counts, void* stream) {
    ABCDLS_CHECK_ARG(h, h && model_id && idx && n_acc && counts && K >= 1 && K <= 4096 && cap >= 1, "postpr_counts");
    cudaStream_t st = (cudaStream_t)stream;
    ABCDLS_CUDA(h, cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int64_t), st));
    k_postpr_counts<<<grid_for(h, cap, 256, 4), 256, (size_t)K * sizeof(unsigned), st>>>(
        model_id, (const long long*)idx, (const long long*)n_acc, row_offset, cap, K, (unsigned long long*)counts);
    ABCDLS_LAUNCH_CHECK(h);
    return ABCDLS_OK;
}

}  // extern "C"
