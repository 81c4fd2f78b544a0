// single.cu — the call-level entry point of the C ABI: abc(target, param, sumstat, tol, method) on ONE GPU as one C
// call, for hosts that do not want to mirror the stage sequencing of abc-dls_b200/engine.py + stages.py (SURVEY §8b
// sketched coarse calls; VERDICT r1 "a non-Python host cannot use the engine without re-writing that").
//
// Replaces, for one rank and the one-pass front half, what the reference binds through rpy2:
//   abc.abc(target=, param=, sumstat=, tol=, method=)      /root/reference/src/Classes/ABC.py:1950-1953, 1964, 2806-2809
// It only SEQUENCES the stage-level entry points (same kernels, same order as AbcEngine._front / _abc_device with
// world == 1): sample_cols [-> sample_fine] -> sample_dist -> pass -> rows -> refine x 4 -> exact -> ds x 3 -> check ->
// count_le -> accept -> tail (gather_fit [-> resid_fit] -> adjust_final).  Everything is enqueued on `stream`; nothing
// synchronises.  Data-dependent outcomes (R's stop() conditions, ABCDLS_S_FASTPATH_MISS = "the one-pass path could not
// prove its result complete: rerun through the two-pass / radix stages") come back in small_out[0] like everywhere else.
#include "common.cuh"

#include <math.h>

namespace {

constexpr int kSingleFineShift = 5;                 // second-stage sample: one line in 32 ...
constexpr long long kSingleFineMinRows = 1ll << 23; // ... from 2^23 rows (the engine's policy)

struct SingleWs {
    int* status;
    long long* state;      // n_acc | le | n_cand | (unused)
    long long* quota;
    double* med;
    double* mad;
    double* ds;
    long long* cand_row;
    double* cand_dist;
    long long* perm;
    long long* slot;
    int* cand_lrow;
    double* cand_ss;
    void* fused;
    void* accept;
    void* tail;
    size_t fused_bytes, accept_bytes, tail_bytes, head_bytes, total;
    long long nacc, cap, ccap, ecap;
    int pitch;
};

inline SingleWs carve_single(abcdls_handle_t h, void* ws, long long n, int d, int P, double tol) {
    SingleWs w = {};
    w.nacc = (long long)ceil((double)n * tol);
    w.cap = w.nacc < n ? w.nacc : n;
    w.ccap = 2 * w.nacc + 65536;
    if (w.ccap > n) w.ccap = n;
    w.ecap = abcdls_fused_emit_cap(w.ccap);
    w.pitch = abcdls_fused_emit_pitch(d);
    w.fused_bytes = abcdls_fused_workspace_bytes(n, d);
    w.accept_bytes = abcdls_accept_workspace_bytes(w.ccap);
    w.tail_bytes = abcdls_tail_workspace_bytes(h, d, P);
    char* p = reinterpret_cast<char*>(ws);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t a = o; o += align_up(bytes ? bytes : 8, 256); return p ? p + a : (char*)nullptr; };
    w.status = (int*)take(8);
    w.state = (long long*)take(4 * 8);
    w.quota = (long long*)take(8);
    w.head_bytes = o;                                // zeroed / initialised at the start of every call
    w.med = (double*)take((size_t)d * 8);
    w.mad = (double*)take((size_t)d * 8);
    w.ds = (double*)take(8);
    w.cand_row = (long long*)take((size_t)w.ccap * 8);
    w.cand_dist = (double*)take((size_t)w.ccap * 8);
    w.perm = (long long*)take((size_t)w.ccap * 8);
    w.slot = (long long*)take((size_t)w.cap * 8);
    w.cand_lrow = (int*)take((size_t)w.ecap * 4);
    w.cand_ss = (double*)take((size_t)w.ecap * w.pitch * 8);
    w.fused = take(w.fused_bytes);
    w.accept = take(w.accept_bytes);
    w.tail = take(w.tail_bytes);
    w.total = o;
    return w;
}

__global__ void k_single_init(int* status, long long* state, long long* quota, long long nacc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        *status = 0;
        state[0] = state[1] = state[2] = state[3] = 0;
        *quota = nacc;
    }
}

bool single_ok(int64_t n, int d, int P, double tol) {
    return n >= 1 && d >= 1 && P >= 1 && tol > 0.0 && tol <= 1.0 && abcdls_tail_supported(d, P) &&
           (long long)ceil((double)n * tol) >= 1;
}

}  // namespace

extern "C" {

size_t abcdls_abc_single_workspace_bytes(abcdls_handle_t h, int64_t n, int d, int P, double tol) {
    if (!h || !single_ok(n, d, P, tol)) return 0;
    return carve_single(h, nullptr, n, d, P, tol).total;
}

int64_t abcdls_abc_single_nacc(int64_t n, double tol) { return (int64_t)ceil((double)n * tol); }

#define SINGLE_TRY(call)        \
    do {                        \
        int _rc = (call);       \
        if (_rc != ABCDLS_OK) return _rc; \
    } while (0)

int abcdls_abc_single(abcdls_handle_t h, const double* target, const double* param_cm, int64_t ld_param,
                      const double* param_rm, int64_t row_pitch, const double* sumstat, int64_t ld, int64_t n, int d,
                      int P, double tol, int loclinear, int hcorr, int64_t* idx_out, double* dist_out, double* y,
                      double* ss_out, double* xs, double* w, double* adj, double* res, double* small_out, void* ws,
                      size_t ws_bytes, void* stream) {
    ABCDLS_CHECK_ARG(h, h && target && sumstat && idx_out && dist_out && y && ss_out && small_out && ws,
                     "abc_single: null pointer");
    ABCDLS_CHECK_ARG(h, (param_cm != nullptr) != (param_rm != nullptr), "abc_single: one parameter table");
    ABCDLS_CHECK_ARG(h, single_ok(n, d, P, tol), "abc_single: shape / tol");
    ABCDLS_CHECK_ARG(h, !loclinear || (xs && w && adj && res), "abc_single: loclinear needs xs, w, adj, res");
    ABCDLS_CHECK_ARG(h, abcdls_fused_supported(sumstat, n, d, ld),
                     "abc_single: the table does not qualify for the one-pass path (d <= 32, n >= 16384, even pitch, "
                     "16-byte aligned): use the stage-level entry points");
    SingleWs W = carve_single(h, ws, n, d, P, tol);
    if (ws_bytes < W.total) {
        h->err = "abc_single workspace too small";
        return ABCDLS_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int fine = n >= kSingleFineMinRows ? kSingleFineShift : 0;
    const int lin = loclinear ? 1 : 0, hc = (lin && hcorr) ? 1 : 0;
    long long* n_acc = W.state + 0;
    long long* le = W.state + 1;
    long long* n_cand = W.state + 2;

    k_single_init<<<1, 32, 0, st>>>(W.status, W.state, W.quota, W.nacc);
    ABCDLS_LAUNCH_CHECK(h);
    ABCDLS_CUDA(h, cudaMemsetAsync(W.tail, 0, W.tail_bytes, st));        // the tail kernels' ticket counters

    // ---- front half: medians, MADs and the candidate rows of the cut from ONE table read ----
    SINGLE_TRY(abcdls_fused_sample_cols(h, sumstat, n, d, ld, 1, fine, W.fused, W.fused_bytes, nullptr, nullptr, stream));
    if (fine)
        SINGLE_TRY(abcdls_fused_sample_fine(h, sumstat, n, d, ld, 1, fine, W.fused, W.fused_bytes, nullptr, nullptr, stream));
    SINGLE_TRY(abcdls_fused_sample_dist(h, sumstat, n, d, ld, target, W.nacc, n, 1, fine, W.fused, W.fused_bytes, nullptr,
                                        nullptr, stream));
    SINGLE_TRY(abcdls_fused_pass(h, sumstat, n, d, ld, 1, fine, W.cand_ss, W.cand_lrow, W.ecap, W.fused, W.fused_bytes,
                                 W.status, nullptr, nullptr, stream));
    SINGLE_TRY(abcdls_fused_rows(h, sumstat, n, d, ld, 1, 0, W.ccap, (int64_t*)W.cand_row, (int64_t*)n_cand, W.fused,
                                 W.fused_bytes, W.status, stream));
    for (int phase = 0; phase < 4; ++phase)
        SINGLE_TRY(abcdls_fused_refine(h, sumstat, n, n, d, ld, phase, 1, nullptr, nullptr, W.fused, W.fused_bytes, W.med,
                                       W.mad, W.status, nullptr, nullptr, nullptr, nullptr, stream));
    SINGLE_TRY(abcdls_fused_exact(h, sumstat, n, d, ld, 1, W.mad, target, W.ccap, W.cand_dist, W.cand_ss, W.cand_lrow,
                                  W.ecap, (int64_t*)W.perm, W.fused, W.fused_bytes, W.status, stream));
    for (int phase = 0; phase < 3; ++phase) {
        void* x0 = nullptr;
        size_t c0 = 0;
        SINGLE_TRY(abcdls_fused_ds(h, n, d, phase, 1, nullptr, W.cand_dist, (const int64_t*)n_cand, W.ccap, W.nacc, W.ds,
                                   W.fused, W.fused_bytes, W.status, &x0, &c0, stream));
    }
    SINGLE_TRY(abcdls_fused_check(h, n, d, W.mad, W.ds, W.fused, W.fused_bytes, W.status, stream));

    // ---- R's cap rule on the candidate list: the first nacc rows in row order among dist <= ds ----
    SINGLE_TRY(abcdls_count_le(h, W.cand_dist, W.ccap, (const int64_t*)n_cand, W.ds, (int64_t*)le, W.accept,
                               W.accept_bytes, stream));
    SINGLE_TRY(abcdls_accept(h, W.cand_dist, W.ccap, (const int64_t*)n_cand, (const int64_t*)W.cand_row, W.ds,
                             (const int64_t*)W.quota, 0, W.cap, idx_out, dist_out, (int64_t*)W.slot, (int64_t*)n_acc,
                             W.accept, W.accept_bytes, W.status, stream));

    // ---- the tail: accepted rows -> posterior sample (+ every small result in small_out) ----
    SINGLE_TRY(abcdls_tail_gather_fit(h, nullptr, W.cand_ss, W.pitch, (const int64_t*)W.slot, (const int64_t*)W.perm,
                                      nullptr, 0, param_cm, ld_param, param_rm, row_pitch, dist_out, idx_out,
                                      (const int64_t*)n_acc, 0, W.cap, d, P, W.mad, target, W.ds, lin, xs, y, ss_out, w,
                                      W.tail, W.tail_bytes, W.status, stream));
    if (hc)
        SINGLE_TRY(abcdls_tail_resid_fit(h, nullptr, xs, y, w, (const int64_t*)n_acc, W.cap, d, P, W.tail, W.tail_bytes,
                                         W.status, stream));
    SINGLE_TRY(abcdls_tail_adjust_final(h, nullptr, xs, y, w, (const int64_t*)n_acc, W.ds, W.cap, d, P, lin, hc, W.mad,
                                        target, adj, res, W.tail, W.tail_bytes, small_out, W.status, stream));
    return ABCDLS_OK;
}

}  // extern "C"
