// peer.cuh — NVLink peer-memory mailboxes: the device-side protocol shared by the stand-alone exchange kernels
// (peer.cu) and by the kernels that embed an exchange in their last thread block (tail.cu: normal equations ->
// all-reduce -> Cholesky in ONE launch).  See peer.cu for the protocol description.
#pragma once
#include "common.cuh"

constexpr int kMaxWorld = 64;
constexpr int kPeerThreads = 512;
constexpr int kPeerMaxCtas = 32;
// A peer that never publishes (crashed rank, a rank that raised before enqueueing its half of the call) must not hang
// this GPU inside a kernel that cannot be cancelled: every wait gives up after PeerDev::spin_ns and raises the sticky
// error word, which the end-of-call kernel turns into ABCDLS_S_PEER_TIMEOUT.  Default 60 s (ABCDLS_PEER_TIMEOUT_S at
// abcdls_peer_create): a rank may legitimately be seconds late (host work between calls - the 4 s of the first version
// fired in the 8-GPU parity test while rank 0 was still computing the previous case's CPU oracle).
constexpr unsigned long long kPeerSpinNsDefault = 60000000000ull;

struct PeerDev {
    unsigned long long* seq;       // exchanges completed by this rank (device memory, advanced by the kernels)
    unsigned long long* bar;       // arrivals of my CTAs (monotonic)
    unsigned long long* bar_cum;   // arrivals expected before the current exchange
    unsigned long long* err;       // sticky: != 0 after a wait timed out
    unsigned long long spin_ns;    // budget of every wait
    unsigned long long* trace;     // NULL, or kPeerTraceCap x 4 timestamps (ABCDLS_PEER_TRACE=1): see peer_trace
    unsigned char* box[kMaxWorld]; // mailboxes of all ranks as mapped in THIS process (box[rank] = mine)
    size_t slot_bytes;
    int rank, world;
};

struct abcdls_peer_s {
    PeerDev dev;
    void* mine = nullptr;
    void* opened[kMaxWorld] = {};
    unsigned long long* ctl = nullptr;   // seq | bar | bar_cum | err
    unsigned long long* trace = nullptr;
    size_t box_bytes = 0;
    std::string err;
};

// Optional timeline of the exchanges (ABCDLS_PEER_TRACE=1): record k % kPeerTraceCap holds %globaltimer at
// {entry, own contribution published, every peer's flag seen, exit} of exchange k, written by one thread.
constexpr int kPeerTraceCap = 4096;
__device__ __forceinline__ void peer_trace(const PeerDev& P, unsigned long long k, int what) {
    if (P.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        P.trace[(k % kPeerTraceCap) * 4 + what] = t;
    }
}

constexpr size_t kFlagsBytes = (size_t)2 * kMaxWorld * sizeof(unsigned long long);   // 1024
static_assert(kFlagsBytes % 256 == 0, "data slots start 256-byte aligned");

__device__ __forceinline__ unsigned long long* flag_ptr(unsigned char* box, int parity, int src) {
    return reinterpret_cast<unsigned long long*>(box) + (size_t)parity * kMaxWorld + src;
}
__device__ __forceinline__ unsigned char* data_ptr(unsigned char* box, size_t slot_bytes, int parity) {
    return box + kFlagsBytes + (size_t)parity * slot_bytes;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want; false (and the sticky error word set) when the budget runs out
__device__ __forceinline__ bool spin_until_ge(const unsigned long long* p, unsigned long long want,
                                              unsigned long long* err, unsigned long long budget_ns) {
    if (ld_volatile_u64(p) >= want) return true;
    const unsigned long long t0 = global_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i) {
            if (ld_volatile_u64(p) >= want) return true;
            __nanosleep(20);
        }
        if (ld_volatile_u64(err) != 0ull) return false;          // somebody already gave up: do not wait 4 s each
        if (global_ns() - t0 > budget_ns) {
            atomicExch(err, 1ull);
            return false;
        }
    }
}

// ---- an exchange embedded in ONE thread block of a compute kernel -------------------------------------------------
// vec[0 .. count): the first n_min entries are reduced with min, the rest with sum, over all ranks in RANK ORDER
// (bit-identical on every rank).  All threads of the block call it; vec may be shared or global memory written by this
// block.  Same sequence numbers / slot parities as the stand-alone kernels, so both kinds interleave on one stream.
__device__ __forceinline__ void peer_block_allreduce(const PeerDev& P, double* vec, int count, int n_min) {
    const unsigned long long k = *P.seq + 1ull;
    const int parity = (int)(k & 1ull);
    double* mine = reinterpret_cast<double*>(data_ptr(P.box[P.rank], P.slot_bytes, parity));
    __syncthreads();
    if (threadIdx.x == 0) peer_trace(P, k, 0);
    for (int i = threadIdx.x; i < count; i += blockDim.x) mine[i] = vec[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) peer_trace(P, k, 1);
    if (threadIdx.x < P.world) {
        unsigned long long* f = flag_ptr(P.box[threadIdx.x], parity, P.rank);   // my entry in peer threadIdx.x's row
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(f), "l"(k) : "memory");
        spin_until_ge(flag_ptr(P.box[P.rank], parity, threadIdx.x), k, P.err, P.spin_ns);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) peer_trace(P, k, 2);
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        double acc = 0.0;
        for (int r0 = 0; r0 < P.world; r0 += 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < P.world)
                    v[q] = __ldcv(reinterpret_cast<const double*>(data_ptr(P.box[r0 + q], P.slot_bytes, parity)) + i);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < P.world) {
                    const double x = v[q];
                    acc = (r0 + q == 0) ? x : (i < n_min ? (x < acc ? x : acc) : acc + x);
                }
        }
        vec[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        peer_trace(P, k, 3);
        *P.seq = k;
    }
    __syncthreads();
}
