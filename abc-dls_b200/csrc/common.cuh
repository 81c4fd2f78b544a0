// common.cuh — shared helpers for the abcdls sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

#include "../../include/abcdls.h"

struct abcdls_handle_s {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    int no_tma = 0;   // ABCDLS_NO_TMA=1 in the environment: use the load/store distance kernel (A/B comparisons)
    std::string err;
};

#define ABCDLS_CHECK_ARG(h, cond, msg)                \
    do {                                              \
        if (!(cond)) {                                \
            if (h) (h)->err = std::string("invalid argument: ") + (msg); \
            return ABCDLS_E_INVALID;                  \
        }                                             \
    } while (0)

#define ABCDLS_CUDA(h, call)                                                          \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            if (h) (h)->err = std::string(#call) + ": " + cudaGetErrorString(_e);     \
            return ABCDLS_E_CUDA;                                                     \
        }                                                                             \
    } while (0)

#define ABCDLS_LAUNCH_CHECK(h)                                                        \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            if (h) (h)->err = std::string("kernel launch: ") + cudaGetErrorString(_e); \
            return ABCDLS_E_CUDA;                                                     \
        }                                                                             \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// A call is a chain of ~45 mostly tiny, dependent kernels; between two of them the GPU idles for the launch latency
// (~2-3 us in a CUDA graph).  Kernels of the hot path start with pdl_prologue(): "my dependents may be scheduled now"
// + "wait until everything before me has completed and is visible", and are launched with the programmatic stream
// serialization attribute, so the NEXT kernel's blocks are already resident (parked at the wait) when this one ends.
// Launched without the attribute (ABCDLS_NO_PDL=1, or the other paths) both instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("ABCDLS_NO_PDL");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on != 0;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- order-preserving 64-bit key of a double -------------------------------------------------
__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double keyd(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---- streaming loads (read-once data: keep it out of L1) -------------------------------------
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// ---- block reductions ------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exclusive scan of one int per thread across the block (blockDim.x <= 1024, multiple of 32).
// returns the exclusive prefix; *total gets the block total.  `sh` must hold 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* sh, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? sh[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        sh[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    int res = sh[wid] + incl - v;
    *total = sh[32];
    __syncthreads();
    return res;
}

// Correctly rounded a / b for a fixed divisor b, given rb = RN(1/b):
//   q0 = RN(a*rb); r0 = a - b*q0 (exact, fma); q1 = RN(q0 + r0*rb)  -> faithful
//   r1 = a - b*q1 (exact);             q2 = RN(q1 + r1*rb)           -> RN(a/b)  (Markstein)
// Valid while every intermediate stays in the normal range; callers guard with |a| in (lo, hi)
// and fall back to IEEE division outside (a == 0 is exact on the fast path).
__device__ __forceinline__ double div_const(double a, double b, double rb) {
    double q0 = __dmul_rn(a, rb);
    double r0 = __fma_rn(-b, q0, a);
    double q1 = __fma_rn(r0, rb, q0);
    double r1 = __fma_rn(-b, q1, a);
    return __fma_rn(r1, rb, q1);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) -------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, `bytes` multiple of 16, both addresses 16-byte aligned; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

