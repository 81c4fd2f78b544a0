// peer.cu — latency-bound collectives over NVLink peer memory (no NCCL on the data path).
//
// The ABC path shards by rows and exchanges only small messages between its stages (sample estimates, band counts,
// histograms, the (d+1) x (d+1+P) normal-equation partials, <= 1 MB refinement blocks).  NCCL costs
// 30-60 us of launch + protocol latency for each of the ~25 exchanges of a call and cannot live inside our CUDA
// graph.  Here every rank owns a "mailbox" (cudaMalloc + CUDA IPC, mapped by all peers over NVLink / NVSwitch):
//
//     flags[2][world]   sequence numbers PUSHED by the peers (every rank polls its own memory)
//     data [2][slot]    what this rank contributes to the current exchange (double buffered by sequence parity)
//
// One kernel per exchange: copy my contribution into my data slot -> fence -> push the sequence number into every
// peer's flag row -> wait for all flags -> read the peers' slots over NVLink and reduce IN RANK ORDER (every rank
// computes bit-identical results; deterministic) or concatenate (all-gather).  The sequence counter lives in device
// memory and is advanced by the kernel itself, so the kernels are ordinary graph nodes: a multi-rank call replays as
// ONE CUDA graph per rank.  Slot reuse is safe with two parities: a rank can only start exchange k+2 after every
// peer has published k+1, i.e. after every peer has finished reading exchange k.
#include "peer.cuh"

#include <string.h>

namespace {

inline size_t flags_bytes() { return kFlagsBytes; }

// all CTAs of this launch have arrived (monotonic counter; *bar_cum = arrivals before this exchange)
__device__ __forceinline__ void grid_arrive_wait(const PeerDev& P) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        atomicAdd(P.bar, 1ull);
        spin_until_ge(P.bar, *P.bar_cum + gridDim.x, P.err, P.spin_ns);
    }
    __syncthreads();
}

// publish my contribution (already written to my data slot by all CTAs) and wait for everybody's
__device__ __forceinline__ void publish_and_wait(const PeerDev& P, unsigned long long k, int parity) {
    if (blockIdx.x == 0 && threadIdx.x == 0) peer_trace(P, k, 0);
    grid_arrive_wait(P);   // every CTA's part of the slot is written and fenced
    if (blockIdx.x == 0 && threadIdx.x == 0) peer_trace(P, k, 1);
    if (blockIdx.x == 0 && threadIdx.x < P.world) {
        unsigned long long* f = flag_ptr(P.box[threadIdx.x], parity, P.rank);   // my entry in peer threadIdx.x's row
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(f), "l"(k) : "memory");
    }
    if (threadIdx.x < P.world) spin_until_ge(flag_ptr(P.box[P.rank], parity, threadIdx.x), k, P.err, P.spin_ns);
    __threadfence_system();
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) peer_trace(P, k, 2);
}

__device__ __forceinline__ void finish(const PeerDev& P, unsigned long long k) {
    // the LAST CTA to get here advances the sequence number and the barrier base for the next exchange
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long a = atomicAdd(P.bar, 1ull) + 1ull;
        if (a == *P.bar_cum + 2ull * gridDim.x) {
            peer_trace(P, k, 3);
            *P.bar_cum = a;
            *P.seq = k;
        }
    }
}

// op: 0 sum, 1 min, 2 max;  T = double | long long.
// The peers' slots are read with ld.global.cv (never from a stale L1 line; the slot addresses repeat every second
// exchange) as ordinary loads the compiler may batch: all peers' values of an element are in flight together, so an
// exchange costs ONE NVLink round trip, not world - 1.
template <typename T>
__global__ void __launch_bounds__(kPeerThreads) k_peer_allreduce(PeerDev P, T* __restrict__ data, size_t count, int op) {
    pdl_prologue();
    const unsigned long long k = *P.seq + 1ull;
    const int parity = (int)(k & 1ull);
    T* mine = reinterpret_cast<T*>(data_ptr(P.box[P.rank], P.slot_bytes, parity));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < count; i += stride) mine[i] = data[i];
    publish_and_wait(P, k, parity);
    for (size_t i = i0; i < count; i += stride) {
        T acc = T(0);
        for (int r0 = 0; r0 < P.world; r0 += 8) {   // rank order: identical bits on every rank
            T v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < P.world)
                    v[q] = __ldcv(reinterpret_cast<const T*>(data_ptr(P.box[r0 + q], P.slot_bytes, parity)) + i);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < P.world) {
                    const T x = v[q];
                    acc = (r0 + q == 0) ? x : ((op == 0) ? acc + x : (op == 1 ? (x < acc ? x : acc) : (x > acc ? x : acc)));
                }
        }
        data[i] = acc;
    }
    finish(P, k);
}

// dst[r * words .. ) = contribution of rank r (8-byte words)
__global__ void __launch_bounds__(kPeerThreads) k_peer_allgather(PeerDev P, const long long* __restrict__ src, size_t words,
                                                                 long long* __restrict__ dst) {
    pdl_prologue();
    const unsigned long long k = *P.seq + 1ull;
    const int parity = (int)(k & 1ull);
    long long* mine = reinterpret_cast<long long*>(data_ptr(P.box[P.rank], P.slot_bytes, parity));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < words; i += stride) mine[i] = src[i];
    publish_and_wait(P, k, parity);
    for (size_t i = i0; i < words; i += stride) {
        for (int r0 = 0; r0 < P.world; r0 += 8) {
            long long v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < P.world)
                    v[q] = __ldcv(reinterpret_cast<const long long*>(data_ptr(P.box[r0 + q], P.slot_bytes, parity)) + i);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (r0 + q < P.world) dst[(size_t)(r0 + q) * words + i] = v[q];
        }
    }
    finish(P, k);
}

// nj ragged rows per rank: src_vals[nj][cap] of which only the first src_cnt[j] entries of row j exist (the refinement
// blocks: <= 8192 band values per column, a few hundred in practice).  dst_vals[world][nj][cap] / dst_cnt[world][nj]
// get every rank's rows, but only the existing entries cross NVLink (the padded block is 1 MB per rank at d = 16).
// Slot layout: nj counts (8-byte words) | nj x cap values.  CTA b serves the (rank, row) pairs b, b + grid, ...
__global__ void __launch_bounds__(kPeerThreads) k_peer_allgather_ragged(PeerDev P, const double* __restrict__ src_vals,
                                                                        const unsigned* __restrict__ src_cnt, int nj,
                                                                        long long cap, double* __restrict__ dst_vals,
                                                                        unsigned* __restrict__ dst_cnt) {
    pdl_prologue();
    const unsigned long long k = *P.seq + 1ull;
    const int parity = (int)(k & 1ull);
    unsigned long long* mine = reinterpret_cast<unsigned long long*>(data_ptr(P.box[P.rank], P.slot_bytes, parity));
    double* mine_v = reinterpret_cast<double*>(mine + nj);
    for (int j = blockIdx.x; j < nj; j += gridDim.x) {
        long long c = src_cnt[j];
        if (threadIdx.x == 0) mine[j] = (unsigned long long)c;     // the raw count: an overflowed row must stay visible
        if (c > cap) c = cap;
        for (long long i = threadIdx.x; i < c; i += blockDim.x) mine_v[(size_t)j * cap + i] = src_vals[(size_t)j * cap + i];
    }
    publish_and_wait(P, k, parity);
    for (int pair = blockIdx.x; pair < P.world * nj; pair += gridDim.x) {
        const int r = pair / nj, j = pair - r * nj;
        const unsigned long long* peer = reinterpret_cast<const unsigned long long*>(data_ptr(P.box[r], P.slot_bytes, parity));
        const double* pv = reinterpret_cast<const double*>(peer + nj) + (size_t)j * cap;
        double* out = dst_vals + ((size_t)r * nj + j) * cap;
        // the count and the first blockDim.x values travel together: ONE NVLink round trip for a typical row (a few
        // hundred entries) instead of two dependent ones
        const double v0 = threadIdx.x < cap ? __ldcv(pv + threadIdx.x) : 0.0;
        long long c = (long long)__ldcv(peer + j);
        if (threadIdx.x == 0) dst_cnt[(size_t)r * nj + j] = (unsigned)c;
        if (c > cap) c = cap;
        if (threadIdx.x < c) out[threadIdx.x] = v0;
        for (long long i = threadIdx.x + blockDim.x; i < c; i += blockDim.x) out[i] = __ldcv(pv + i);
    }
    finish(P, k);
}

}  // namespace

extern "C" {

size_t abcdls_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

// Allocates this rank's mailbox and writes its CUDA IPC handle (abcdls_peer_handle_bytes() bytes) to handle_out.
int abcdls_peer_create(abcdls_peer_t* out, int rank, int world, size_t slot_bytes, void* handle_out) {
    if (!out || !handle_out || world < 2 || world > kMaxWorld || rank < 0 || rank >= world || slot_bytes < 1024)
        return ABCDLS_E_INVALID;
    abcdls_peer_s* p = new abcdls_peer_s();
    slot_bytes = align_up(slot_bytes, 256);
    p->box_bytes = flags_bytes() + 2 * slot_bytes;
    if (cudaMalloc(&p->mine, p->box_bytes) != cudaSuccess || cudaMalloc((void**)&p->ctl, 256) != cudaSuccess) {
        cudaGetLastError();
        delete p;
        return ABCDLS_E_CUDA;
    }
    cudaMemset(p->mine, 0, p->box_bytes);
    cudaMemset(p->ctl, 0, 256);
    cudaIpcMemHandle_t hnd;
    if (cudaIpcGetMemHandle(&hnd, p->mine) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p->mine);
        cudaFree(p->ctl);
        delete p;
        return ABCDLS_E_CUDA;
    }
    memcpy(handle_out, &hnd, sizeof(hnd));
    p->dev.seq = p->ctl;
    p->dev.bar = p->ctl + 8;
    p->dev.bar_cum = p->ctl + 16;
    p->dev.err = p->ctl + 24;
    p->dev.spin_ns = kPeerSpinNsDefault;
    if (const char* e = getenv("ABCDLS_PEER_TIMEOUT_S")) {
        const double sec = atof(e);
        if (sec > 0.0) p->dev.spin_ns = (unsigned long long)(sec * 1e9);
    }
    p->dev.trace = nullptr;
    if (const char* e = getenv("ABCDLS_PEER_TRACE"))
        if (e[0] == '1' && cudaMalloc((void**)&p->trace, (size_t)kPeerTraceCap * 32) == cudaSuccess) {
            cudaMemset(p->trace, 0, (size_t)kPeerTraceCap * 32);
            p->dev.trace = p->trace;
        }
    p->dev.slot_bytes = slot_bytes;
    p->dev.rank = rank;
    p->dev.world = world;
    for (int r = 0; r < kMaxWorld; ++r) p->dev.box[r] = nullptr;
    p->dev.box[rank] = reinterpret_cast<unsigned char*>(p->mine);
    cudaDeviceSynchronize();
    *out = p;
    return ABCDLS_OK;
}

// all_handles: world x abcdls_peer_handle_bytes() bytes, rank order (gathered by the caller's bootstrap channel)
int abcdls_peer_connect(abcdls_peer_t p, const void* all_handles) {
    if (!p || !all_handles) return ABCDLS_E_INVALID;
    const char* hb = reinterpret_cast<const char*>(all_handles);
    for (int r = 0; r < p->dev.world; ++r) {
        if (r == p->dev.rank) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, hb + (size_t)r * sizeof(hnd), sizeof(hnd));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            p->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
            cudaGetLastError();
            return ABCDLS_E_CUDA;
        }
        p->opened[r] = ptr;
        p->dev.box[r] = reinterpret_cast<unsigned char*>(ptr);
    }
    return ABCDLS_OK;
}

int abcdls_peer_destroy(abcdls_peer_t p) {
    if (!p) return ABCDLS_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < kMaxWorld; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
    if (p->mine) cudaFree(p->mine);
    if (p->ctl) cudaFree(p->ctl);
    if (p->trace) cudaFree(p->trace);
    delete p;
    return ABCDLS_OK;
}

// ABCDLS_PEER_TRACE=1: copies the timeline (kPeerTraceCap records of 4 x uint64 ns: entry, published, flags seen, exit;
// record k % cap = exchange k) and the number of exchanges so far to the host.  Returns the record capacity (0: off).
int abcdls_peer_trace_read(abcdls_peer_t p, uint64_t* out, uint64_t* n_exchanges) {
    if (!p || !p->trace || !out) return 0;
    cudaDeviceSynchronize();
    cudaMemcpy(out, p->trace, (size_t)kPeerTraceCap * 32, cudaMemcpyDeviceToHost);
    if (n_exchanges) cudaMemcpy(n_exchanges, p->dev.seq, 8, cudaMemcpyDeviceToHost);
    return kPeerTraceCap;
}

const char* abcdls_peer_last_error(abcdls_peer_t p) { return p ? p->err.c_str() : "null peer"; }

size_t abcdls_peer_slot_bytes(abcdls_peer_t p) { return p ? p->dev.slot_bytes : 0; }

static inline unsigned peer_grid(size_t bytes) {
    size_t g = (bytes + 4 * 1024 - 1) / (4 * 1024);   // one 8-byte element per thread up to 32 CTAs
    if (g < 1) g = 1;
    if (g > (size_t)kPeerMaxCtas) g = kPeerMaxCtas;
    return (unsigned)g;
}

// in place; dtype: 0 = double, 1 = int64; op: 0 sum, 1 min, 2 max.  count * 8 <= slot bytes.
int abcdls_peer_allreduce(abcdls_peer_t p, void* data, size_t count, int dtype, int op, void* stream) {
    if (!p || !data || count * 8 > p->dev.slot_bytes || dtype < 0 || dtype > 1 || op < 0 || op > 2) return ABCDLS_E_INVALID;
    if (count == 0) return ABCDLS_OK;
    const unsigned g = peer_grid(count * 8);
    if (dtype == 0)
        launch_pdl(k_peer_allreduce<double>, dim3(g), dim3(kPeerThreads), 0, (cudaStream_t)stream, p->dev, (double*)data, count, op);
    else
        launch_pdl(k_peer_allreduce<long long>, dim3(g), dim3(kPeerThreads), 0, (cudaStream_t)stream, p->dev, (long long*)data, count, op);
    return cudaGetLastError() == cudaSuccess ? ABCDLS_OK : ABCDLS_E_CUDA;
}

// dst (world * bytes) <- every rank's src (bytes, multiple of 8, <= slot bytes), rank order
int abcdls_peer_allgather(abcdls_peer_t p, const void* src, size_t bytes, void* dst, void* stream) {
    if (!p || !src || !dst || (bytes & 7) || bytes > p->dev.slot_bytes) return ABCDLS_E_INVALID;
    if (bytes == 0) return ABCDLS_OK;
    launch_pdl(k_peer_allgather, dim3(peer_grid(bytes)), dim3(kPeerThreads), 0, (cudaStream_t)stream, p->dev, (const long long*)src, bytes / 8,
                                                                                  (long long*)dst);
    return cudaGetLastError() == cudaSuccess ? ABCDLS_OK : ABCDLS_E_CUDA;
}

// every rank's nj ragged rows (see k_peer_allgather_ragged); (nj + nj * cap) * 8 <= slot bytes
int abcdls_peer_allgather_ragged(abcdls_peer_t p, const double* src_vals, const uint32_t* src_cnt, int nj, int64_t cap,
                                 double* dst_vals, uint32_t* dst_cnt, void* stream) {
    if (!p || !src_vals || !src_cnt || !dst_vals || !dst_cnt || nj < 1 || cap < 1 ||
        ((size_t)nj + (size_t)nj * cap) * 8 > p->dev.slot_bytes)
        return ABCDLS_E_INVALID;
    unsigned g = (unsigned)(p->dev.world * nj);             // one (rank, row) pair per CTA where the device allows
    if (g > 128u) g = 128u;
    launch_pdl(k_peer_allgather_ragged, dim3(g), dim3(kPeerThreads), 0, (cudaStream_t)stream, p->dev, src_vals, src_cnt, nj, cap, dst_vals,
                                                                         dst_cnt);
    return cudaGetLastError() == cudaSuccess ? ABCDLS_OK : ABCDLS_E_CUDA;
}

}  // extern "C"
