// fused_wide.cuh — the one-pass table kernel for 16 < d <= 32 statistics (included by fused.cuh; same protocol, same
// outputs as k_fused_pass).  The reference's own models have 19 and 24 parameters (tests/Model.info), i.e. joint calls
// with 19 / 24 predicted statistics: with the 16-column kernel alone they took the two-pass path (two table reads).
//
// What changes against k_fused_pass: a tile is 128 rows x d columns (a ring stage stays <= 32 KB, so the ring keeps 5+
// stages), every consumer warp classifies TWO columns (w and w + 16: four elements per lane and column, two per-lane
// queues of 8 slots), a row of the tile is shared by FOUR lanes (columns q, q + 4, ..; two shuffles), a warp covers 8
// rows (one byte of the row bitmap per warp and tile: the bitmap stays one bit per row in row order, so every later
// stage is unchanged), and an emitted candidate record has 32 doubles.
// Measured (B200, 5e7 rows): d = 32: 3.02 ms per pass (4.2 TB/s, 0.65 of the measured copy bandwidth); d = 24: 3.24 ms;
// d = 19: 3.10 ms - the time goes with the number of 128-row tiles (1.17 us per tile and SM against 0.81 us for a 256-row
// tile of k_fused_pass), not with the bytes: per-tile overhead (barrier, flush checks, ballot, bitmap byte) and two
// columns' worth of per-warp state.  Splitting the columns beyond the 16th between warp pairs (so that no warp carries
// two full columns when d <= 24) did not change it.  Against the two-pass path it replaces at d = 24: whole call
// 4.57 ms instead of 5.76 ms.
#pragma once

constexpr int kWRows = 128;
constexpr int kWPitch = 32;       // doubles per emitted candidate record
constexpr int kWQCap = 8;         // per-lane queue slots per column (flushed when a lane holds more than 4)

struct WideCol {                  // per-warp state of one classified column
    double c, r0, mr1;
    unsigned hiW, q0, pq;
    double* gM;
    double* gU;
    unsigned long long* cntM;
    unsigned long long* cntU;
    long long posM, endM, posU, endU;
    unsigned c_neg, c_in;
};

// the queue flush of k_fused_pass for one column (see there for the chunk-reservation scheme)
__device__ __forceinline__ void wide_flush(WideCol& S, int lane, long long capM, long long capU, int chunkM, int chunkU) {
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    const int fill = (int)((S.pq - S.q0) >> 8);
    const int mx = __reduce_max_sync(0xffffffffu, fill);
    if (mx == 0) return;
    int mine = 0;
    for (int t = 0; t < mx; ++t) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(S.q0 + t * 256));
        if (t < fill) mine += (fabs(__dsub_rn(v, S.c)) <= S.r0) ? 1 : 0x10000;
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    const long long nM = tot & 0xffff, nU = tot >> 16;
    if (nM > S.endM - S.posM) {
        for (long long i = S.posM + lane; i < S.endM; i += 32)
            if (i < capM) S.gM[i] = qnan;
        const long long take = nM > chunkM ? nM : chunkM;
        unsigned long long b = 0ull;
        if (lane == 0) b = atomicAdd(S.cntM, (unsigned long long)take);
        S.posM = (long long)__shfl_sync(0xffffffffu, b, 0);
        S.endM = S.posM + take;
    }
    if (nU > S.endU - S.posU) {
        for (long long i = S.posU + lane; i < S.endU; i += 32)
            if (i < capU) S.gU[i] = qnan;
        const long long take = nU > chunkU ? nU : chunkU;
        unsigned long long b = 0ull;
        if (lane == 0) b = atomicAdd(S.cntU, (unsigned long long)take);
        S.posU = (long long)__shfl_sync(0xffffffffu, b, 0);
        S.endU = S.posU + take;
    }
    double* pM = S.gM + S.posM + ((incl - mine) & 0xffff);
    double* pU = S.gU + S.posU + ((incl - mine) >> 16);
    S.posM += nM;
    S.posU += nU;
    const bool roomy = S.endM <= capM && S.endU <= capU;
    for (int t = 0; t < mx; ++t) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(S.q0 + t * 256));
        if (t < fill) {
            const bool isM = fabs(__dsub_rn(v, S.c)) <= S.r0;
            double* dst = isM ? pM : pU;
            if (roomy || (isM ? (pM - S.gM) < capM : (pU - S.gU) < capU)) *dst = v;
            pM += isM ? 1 : 0;
            pU += isM ? 0 : 1;
        }
    }
    S.pq = S.q0;
    __syncwarp();
}

__device__ __forceinline__ void wide_classify(WideCol& S, double xv) {
    const double sgn = __dsub_rn(xv, S.c);
    const double uu = __dadd_rn(fabs(sgn), S.mr1);   // |x - c| - r1
    const unsigned hu = (unsigned)__double2hiint(uu);
    S.c_neg += (unsigned)__double2hiint(sgn) >> 31;
    S.c_in += hu >> 31;
    const bool cand = (fabs(sgn) <= S.r0) || (hu <= S.hiW);   // M band or U band
    asm volatile(
        "{\n\t"
        ".reg .pred pc;\n\t"
        "setp.ne.u32 pc, %1, 0;\n\t"
        "@pc st.shared.f64 [%0], %2;\n\t"
        "@pc add.u32 %0, %0, 256;\n\t"
        "}"
        : "+r"(S.pq)
        : "r"((unsigned)cand), "d"(xv));
}

__global__ void __launch_bounds__(kFThreads, 1)
k_fused_pass_wide(const __grid_constant__ CUtensorMap tmap, long long n, long long ntiles, int d, int nstages,
                  const ColBrk* __restrict__ brk, const FusedPar* __restrict__ par, unsigned long long* __restrict__ counts,
                  double* __restrict__ candM, long long capM, double* __restrict__ candU, long long capU,
                  unsigned long long* __restrict__ cnt, unsigned char* __restrict__ mask8, double* __restrict__ cand_e,
                  long long ecap, int* __restrict__ erow, unsigned long long* __restrict__ ealloc, long long tpc,
                  int chunkM, int chunkU, int* status) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
    double* ring = reinterpret_cast<double*>(smem_raw + kFHeader);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kFMaxD)
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
    const size_t stage_doubles = (size_t)d * kWRows;
    if (wid == kFWarps) {
        if (lane == 0) {   // producer: ONE 2-D TMA tile copy (128 rows x d columns) per stage
            int st = 0;
            unsigned ph = 0;
            for (long long tile = t_begin; tile < t_end; ++tile) {
                mbar_wait(&sm.empty[st], ph ^ 1u);
                mbar_expect_tx(&sm.full[st], (unsigned)(stage_doubles * 8));
                tma_tile_g2s(ring + (size_t)st * stage_doubles, &tmap, (int)(tile * kWRows), 0, &sm.full[st]);
                if (++st == nstages) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ===== consumers: warp w classifies columns w and w + 16 =====
    unsigned char* qbase = smem_raw + kFHeader + (size_t)nstages * stage_doubles * 8 + (size_t)wid * kFQWarpBytes;
    WideCol col[2];
    bool has[2];
    int jcol[2];
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
        const int j = wid + kFWarps * ci;
        jcol[ci] = j;
        has[ci] = j < d;
        const int jc = has[ci] ? j : 0;
        const ColBrk B = brk[jc];
        WideCol& S = col[ci];
        S.c = B.c;
        S.r0 = B.r0;
        S.mr1 = -B.r1;
        S.hiW = (unsigned)__double2hiint(__dsub_rn(B.r2, B.r1));
        S.q0 = smem_u32(qbase) + (unsigned)ci * (kWQCap * 256) + lane * 8;   // slot s of my queue at q0 + s * 256
        S.pq = S.q0;
        S.gM = candM + (size_t)jc * capM;
        S.gU = candU + (size_t)jc * capU;
        S.cntM = &cnt[jc * 2 + 0];
        S.cntU = &cnt[jc * 2 + 1];
        S.posM = S.endM = S.posU = S.endU = 0;
        S.c_neg = S.c_in = 0;
    }
    const double thr = par->thr;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned eoff = 0;
    int eroom = 0;
    bool bad = false;
    const int r8 = lane & 7, quarter = lane >> 3;
    const int row = (wid << 3) | r8;             // row of the tile this lane helps with
    int st = 0;
    unsigned ph = 0;
    unsigned char* mptr = mask8 + t_begin * kFWarps + wid;
    const long long tail_tile = (n % kWRows) ? ntiles - 1 : -1;
    const int tail_rows = (int)(n % kWRows);
    for (long long tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&sm.full[st], ph);
        const double* __restrict__ stage = ring + (size_t)st * stage_doubles;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
            if (has[ci]) {
                WideCol& S = col[ci];
                if (__any_sync(0xffffffffu, S.pq > S.q0 + (kWQCap - 4) * 256)) wide_flush(S, lane, capM, capU, chunkM, chunkU);
                const double* __restrict__ colp = stage + jcol[ci] * kWRows + 2 * lane;
                const double2 v0 = *reinterpret_cast<const double2*>(colp);
                const double2 v1 = *reinterpret_cast<const double2*>(colp + 64);
                wide_classify(S, v0.x);
                wide_classify(S, v0.y);
                wide_classify(S, v1.x);
                wide_classify(S, v1.y);
            }
        }
        // approximate squared distance of row `row`: this lane sums columns quarter, quarter + 4, ... in two independent
        // chains (the loads and FMAs overlap the classification chains above: the queue is only touched by volatile asm)
        double q = 0.0, qb = 0.0;
        if (d == kFMaxD) {
            const double* __restrict__ xp = stage + quarter * kWRows + row;
            const double2* __restrict__ sp = sm.sc + quarter;
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const double2 s0 = sp[4 * u], s1 = sp[4 * u + 4];
                const double a0 = fma(xp[4 * u * kWRows], s0.y, s0.x);
                const double a1 = fma(xp[(4 * u + 4) * kWRows], s1.y, s1.x);
                q = fma(a0, a0, q);
                qb = fma(a1, a1, qb);
            }
        } else {
            int cc = quarter;
            for (; cc + 4 < d; cc += 8) {
                const double2 s0 = sm.sc[cc], s1 = sm.sc[cc + 4];
                const double a0 = fma(stage[cc * kWRows + row], s0.y, s0.x);
                const double a1 = fma(stage[(cc + 4) * kWRows + row], s1.y, s1.x);
                q = fma(a0, a0, q);
                qb = fma(a1, a1, qb);
            }
            if (cc < d) {
                const double2 s0 = sm.sc[cc];
                const double a0 = fma(stage[cc * kWRows + row], s0.y, s0.x);
                q = fma(a0, a0, q);
            }
        }
        q += qb;
        q += __shfl_xor_sync(0xffffffffu, q, 8);
        q += __shfl_xor_sync(0xffffffffu, q, 16);
        bad |= !(q < inf) && (tile != tail_tile || row < tail_rows);
        const unsigned bal = __ballot_sync(0xffffffffu, q <= thr);
        const unsigned m8 = bal & 0xffu;
        if (lane == 0) *mptr = (unsigned char)m8;
        mptr += kFWarps;
        if (m8) {
            const int k = __popc(m8);
            eroom -= k;
            if (eroom < 0) {
                unsigned nb = 0;
                if (lane == 0) nb = (unsigned)atomicAdd(ealloc, (unsigned long long)kFEChunk);
                eoff = __shfl_sync(0xffffffffu, nb, 0);
                eroom = kFEChunk - k;
            }
            if ((m8 >> r8) & 1u) {
                const unsigned slot = eoff + __popc(m8 & ((1u << r8) - 1u));
                if ((long long)slot < ecap) {
                    // row-major block: 256 bytes per candidate; this lane writes columns [8 quarter, 8 quarter + 8)
                    if (quarter == 0) erow[slot] = (int)(tile * kWRows + row);
                    double* dst = cand_e + (size_t)slot * kWPitch + 8 * quarter;
                    const double* src = stage + (size_t)(8 * quarter) * kWRows + row;
#pragma unroll
                    for (int q2 = 0; q2 < 4; ++q2) {
                        const int c0 = 8 * quarter + 2 * q2;
                        double2 v = make_double2(0.0, 0.0);
                        if (c0 < d) {
                            v.x = src[(2 * q2) * kWRows];
                            if (c0 + 1 < d) v.y = src[(2 * q2 + 1) * kWRows];
                        }
                        *reinterpret_cast<double2*>(dst + 2 * q2) = v;
                    }
                }
            }
            eoff += k;
            __syncwarp();
        }
        if (lane == 0) mbar_arrive(&sm.empty[st]);
        if (++st == nstages) { st = 0; ph ^= 1u; }
    }
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
        if (has[ci]) {
            WideCol& S = col[ci];
            wide_flush(S, lane, capM, capU, chunkM, chunkU);
            for (long long i = S.posM + lane; i < S.endM; i += 32)
                if (i < capM) S.gM[i] = qnan;
            for (long long i = S.posU + lane; i < S.endU; i += 32)
                if (i < capU) S.gU[i] = qnan;
            const unsigned cn = warp_sum(S.c_neg), cin = warp_sum(S.c_in);
            if (lane == 0) {
                const int j = jcol[ci];
                if (cn) atomicAdd(&counts[j * 4 + 0], (unsigned long long)cn);
                if (cin) atomicAdd(&counts[j * 4 + 2], (unsigned long long)cin);
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, ABCDLS_S_NONFINITE);
}
