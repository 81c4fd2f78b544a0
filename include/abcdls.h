/*
 * abcdls.h — C ABI of the B200-native ABC acceptance-and-adjustment engine.
 *
 * This is the drop-in boundary for the step ABC-DLS delegates to R's `abc` package through rpy2
 * (reference: src/Classes/ABC.py:36 binds the package; call sites ABC.py:919, 1950-1953, 1964,
 * 2806-2809, 2817).  The rpy2 FFI for that path is "copy three data frames into R, call
 * abc::abc / abc::postpr, read summary()".  The entry points below are what a native binding for the
 * same path binds instead: plain pointers and sizes, no torch / Python types.
 *
 * Conventions
 *   - every `const double*` / `double*` / `int64_t*` argument is a DEVICE pointer unless its name
 *     ends in `_host`;
 *   - tables are column-major (R `as.matrix` layout): element (i, j) at base[j * ld + i];
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), borrows its inputs
 *     for the duration of the stream work and never modifies them;
 *   - scratch memory is owned by the caller: query the size, allocate, pass it in;
 *   - return value: 0 = ok, otherwise one of ABCDLS_E_*; abcdls_last_error() gives text.
 *     Data-dependent conditions that R reports with stop() (zero variance, ...) cannot be known at
 *     enqueue time; they are OR-ed into the device word `status_dev` (ABCDLS_S_* bits) which the
 *     caller reads back together with the results.
 *   - multi-GPU: rows are sharded in contiguous rank-major blocks.  Stages that need a global
 *     exchange are split so the host can run an NCCL collective between them on the same stream
 *     (the buffers to reduce / gather are returned by the *_buffers calls).
 */
#ifndef ABCDLS_H
#define ABCDLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABCDLS_VERSION_MAJOR 0
#define ABCDLS_VERSION_MINOR 1

/* return codes */
#define ABCDLS_OK 0
#define ABCDLS_E_INVALID 1   /* bad argument (shape, NULL, tol) */
#define ABCDLS_E_WORKSPACE 2 /* workspace too small */
#define ABCDLS_E_CUDA 3      /* CUDA runtime error */
#define ABCDLS_E_NODEVICE 4  /* no sm_100 device */

/* status_dev bits (data-dependent conditions, set by kernels) */
#define ABCDLS_S_NONFINITE 0x1       /* NaN/Inf in sumstat (R: NA rows — not supported) */
#define ABCDLS_S_FASTPATH_MISS 0x2   /* sampled bracket missed / overflowed: rerun with mode EXACT */
#define ABCDLS_S_SINGULAR 0x4        /* normal equations not positive definite (R: lsfit singular) */
#define ABCDLS_S_ZERO_VAR 0x8        /* every sumstat column constant (R stop: Zero variance ...) */
#define ABCDLS_S_ZERO_VAR_REGION 0x10 /* no sumstat column varies inside the accepted region */
#define ABCDLS_S_CAPACITY 0x20       /* an output buffer was too small */
#define ABCDLS_S_PEER_TIMEOUT 0x40   /* a peer never published its part of an exchange (csrc/peer.cuh) */
#define ABCDLS_S_NONFINITE_FIT 0x80  /* hcorr: log(residuals^2) not finite (R: lsfit stops on NA/NaN/Inf in 'y') */

/* selection modes */
#define ABCDLS_MODE_AUTO 0  /* sampled-bracket fast path, verified; sets FASTPATH_MISS on failure */
#define ABCDLS_MODE_EXACT 1 /* distribution-independent 6-pass radix select */

typedef struct abcdls_handle_s* abcdls_handle_t;
typedef struct abcdls_peer_s* abcdls_peer_t; /* NVLink mailbox group, see the end of this file; NULL = one rank */

/* ---- lifetime ---------------------------------------------------------------------------- */
int abcdls_create(abcdls_handle_t* out, int device);
int abcdls_destroy(abcdls_handle_t h);
const char* abcdls_last_error(abcdls_handle_t h); /* host string, valid until the next call on h */
int abcdls_version(void);                         /* major * 100 + minor */
int abcdls_sm_count(abcdls_handle_t h);

/* ---- generic exact selection (order statistics) --------------------------------------------
 * Used for column medians / MADs (replaces R stats::median / stats::mad inside abc:::normalise)
 * and for the tolerance threshold `ds <- sort(dist)[nacc]`.
 *
 * A "selection set" is `njobs` independent problems; job j selects up to two global ranks
 * k0[j] <= k1[j] (0-based) of the values of column j of a column-major table (optionally of
 * |x - center[j]|).  The 64-bit order-preserving key is resolved in ABCDLS_SELECT_PASSES MSB
 * digits; each pass is  hist -> [allreduce of the histogram buffer across ranks] -> pick.
 */
#define ABCDLS_SELECT_PASSES 6
#define ABCDLS_SELECT_BINS 2048

size_t abcdls_select_workspace_bytes(int njobs);
/* job j works on column j of `base` (column-major, leading dimension ld, n rows; counts[j] (device,
 * may be NULL) lets job j use fewer rows).  center (device, may be NULL): select on |x - center[j]|.
 * Ranks: k0_dev/k1_dev (device, per job) or, when NULL, the host scalars k0_host/k1_host for all. */
int abcdls_select_begin(abcdls_handle_t h, void* ws, size_t ws_bytes, int njobs, const double* base,
                        int64_t n, int64_t ld, const int64_t* counts, const double* center,
                        const int64_t* k0_dev, const int64_t* k1_dev, int64_t k0_host, int64_t k1_host,
                        void* stream);
int abcdls_select_hist(abcdls_handle_t h, void* ws, int njobs, int64_t n_max, int pass, void* stream);
int abcdls_select_pick(abcdls_handle_t h, void* ws, int njobs, int pass, void* stream);
/* buffer a multi-GPU caller all-reduces (sum, int64) between hist and pick */
int abcdls_select_hist_buffer(abcdls_handle_t h, void* ws, int njobs, int64_t** buf, size_t* count);
/* results: out[2*j + q] = value of rank kq[j] (the transformed value when center != NULL) */
int abcdls_select_end(abcdls_handle_t h, void* ws, int njobs, double* out, void* stream);
/* R median of an even-length vector = mean of the two middle order statistics:
 * out[j] = scale * (0.5*ranks2[2j] + 0.5*ranks2[2j+1])   (scale 1 -> median, 1.4826 -> mad) */
int abcdls_median_from_ranks(abcdls_handle_t h, const double* ranks2, int n_cols, double scale,
                             double* out, void* stream);

/* ---- kernel 1: MAD scaling + Euclidean distance -------------------------------------------
 * R abc(): normalise() -> sum1 <- sum1 + (scaled[,j]-target[j])^2 -> sqrt   
 */
/* med[j], mad[j] of each column (single-GPU convenience: begin + 6x(hist,pick) + end, twice) */
size_t abcdls_column_mad_workspace_bytes(int d);
int abcdls_column_mad(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, double* med,
                      double* mad, void* ws, size_t ws_bytes, int* status_dev, void* stream);
/* dist[i] = sqrt(sum_j (ss[i,j]/mad[j] - target[j]/mad[j])^2), mad[j]==0 => unscaled.
 * `target` holds the UNSCALED target (device, d doubles).  Sets NONFINITE / ZERO_VAR bits. */
int abcdls_scale_dist(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld,
                      const double* mad, const double* target, double* dist, int* status_dev,
                      void* stream);

/* ---- kernel 2: tolerance cut ---------------------------------------------------------------
 * R: wt1 <- dist <= ds; wt1 <- wt1 & cumsum(wt1) <= nacc    (first nacc rows in row order)
 * `quota` = how many rows this rank may still accept after the lower ranks took theirs
 * (single GPU: nacc).  idx_out gets ascending global row numbers (row_offset + local row).
 */
size_t abcdls_accept_workspace_bytes(int64_t n);
/* count of rows with dist <= ds among the first min(n, *n_dev) entries -> le_count (device int64); feeds the
 * cross-rank scan.  n_dev (device, may be NULL) lets the list length live on the device (candidate lists). */
int abcdls_count_le(abcdls_handle_t h, const double* dist, int64_t n, const int64_t* n_dev, const double* ds,
                    int64_t* le_count, void* ws, size_t ws_bytes, void* stream);
/* must follow abcdls_count_le on the same ws; quota is a device int64.  src_idx (may be NULL): entry i of
 * `dist` belongs to global row src_idx[i] (candidate list) instead of row_offset + i.  dist_out (may be
 * NULL) receives the distances of the accepted rows, slot_out (may be NULL) their positions i in `dist`. */
int abcdls_accept(abcdls_handle_t h, const double* dist, int64_t n, const int64_t* n_dev, const int64_t* src_idx,
                  const double* ds, const int64_t* quota, int64_t row_offset, int64_t cap, int64_t* idx_out,
                  double* dist_out, int64_t* slot_out, int64_t* n_out, void* ws, size_t ws_bytes, int* status_dev,
                  void* stream);

/* ---- fast path: sampled brackets, one table pass each (csrc/bracket.cu) -----------------------
 * Exact results or ABCDLS_S_FASTPATH_MISS in status_dev (then rerun with the radix select above).
 * column median + MAD:  sample -> [all-reduce est] -> pass -> [all-reduce counts] -> finish
 * distance + threshold: sample -> [all-reduce est] -> pass (emits the rows with dist <= bracket top in row
 *                       order, the exact ds, and the N-vector dist only if `dist` != NULL)
 * n must be >= 4096 (smaller tables use the radix path).
 */
size_t abcdls_mad_fast_workspace_bytes(int64_t n, int d);
/* est_out / est_count: buffer (doubles) a multi-rank caller all-reduces (sum) before the pass */
int abcdls_mad_fast_sample(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, void* ws,
                           size_t ws_bytes, double** est_out, size_t* est_count, void* stream);
/* counts_out / counts_count: int64 block (counts | appended | median histogram) to all-reduce (sum) after it */
int abcdls_mad_fast_pass(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, void* ws,
                         size_t ws_bytes, int64_t** counts_out, size_t* counts_count, void* stream);
/* phases 0..3, see csrc/bracket.cu: 0 -> all-gather (x0: doubles, x1: uint32 counts); 1 -> all-reduce x0 (int64);
 * 2 -> all-gather (x0, x1); 3 -> med / mad final.  gath_small / gath_cnt: the gathered block of the previous phase
 * ([world][d][8192] doubles, [world][d] counts), NULL on a single rank. */
int abcdls_mad_fast_refine(abcdls_handle_t h, int64_t n, int64_t n_global, int d, int phase, int world,
                           const double* gath_small, const int32_t* gath_cnt, void* ws, size_t ws_bytes, double* med,
                           double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1, size_t* x1_count,
                           void* stream);
/* single rank: the four phases back to back */
int abcdls_mad_fast_finish(abcdls_handle_t h, int64_t n, int64_t n_global, int d, void* ws, size_t ws_bytes,
                           double* med, double* mad, int* status_dev, void* stream);
size_t abcdls_dist_fast_workspace_bytes(int64_t n, int64_t cap);
int abcdls_dist_fast_sample(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* mad,
                            const double* target, int64_t nacc, int64_t n_global, int world, int64_t cap, void* ws,
                            size_t ws_bytes, double** est_out, size_t* est_count, void* stream);
/* dist may be NULL.  Emits this rank's candidate rows (dist <= bracket top) in row order.  world == 1: ds_out is
 * final on return.  world > 1: all-reduce (sum) the int64 block red_out/red_count, then run the refine phases. */
int abcdls_dist_fast_pass(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* mad,
                          const double* target, int64_t nacc, int world, int64_t row_offset, int64_t cap, double* dist,
                          int64_t* cand_row, double* cand_dist, int64_t* n_cand, double* ds_out, void* ws,
                          size_t ws_bytes, int* status_dev, int64_t** red_out, size_t* red_count, void* stream);
/* phase 0 -> all-gather (x0: 8192 doubles, x1: one uint32 count); phase 1 -> ds_out */
int abcdls_dist_fast_refine(abcdls_handle_t h, int64_t n, int64_t nacc, int64_t cap, int phase, int world,
                            const double* gath_small, const int32_t* gath_cnt, const double* cand_dist, double* ds_out,
                            void* ws, size_t ws_bytes, int* status_dev, void** x0, size_t* x0_count, void** x1,
                            size_t* x1_count, void* stream);

/* ---- one-pass front half (csrc/fused.cuh): medians + MADs AND the tolerance cut from ONE table read ------------
 * R's dependency "MAD before distance" is broken by interval reasoning: sampled MAD estimates with a relative
 * uncertainty eps give an approximate squared distance q~; the pass keeps every row whose q~ could still be
 * <= ds^2, the exact MADs then give those ~1.5 nacc rows R's exact distance, and the device proves that no
 * unmarked row can have dist <= ds (else ABCDLS_S_FASTPATH_MISS: rerun with the two-pass path above).
 *   sample_cols -> [all-reduce est] -> sample_dist -> [all-reduce est] -> pass -> [all-reduce counts]
 *   -> refine phases 0..3 (exchanges as abcdls_mad_fast_refine) -> candidates -> [select ds among the
 *   candidates of all ranks: abcdls_select_*] -> check -> abcdls_count_le / abcdls_accept on the candidate list.
 * Requirements (abcdls_fused_supported): d <= 32 (d <= 16: 256-row tiles, one column per consumer warp; 17..32:
 * 128-row tiles, two columns per warp, csrc/fused_wide.cuh), n >= 16384, ld even, 16-byte aligned table. */
int abcdls_fused_supported(const double* ss, int64_t n, int d, int64_t ld);
size_t abcdls_fused_workspace_bytes(int64_t n, int d);
/* `fine` (the same value in sample_cols / sample_fine / sample_dist / pass of one call, and on every rank): 0 = the
 * sorted first-stage column sample alone brackets median / MAD; s in 1..12 = a second, COUNTING sample of one 128-byte
 * line (16 consecutive rows) out of every 2^s lines of each column refines those brackets (histograms against the
 * first-stage windows; narrower by sqrt(S2 / S1): the pass keeps, and the refinement re-reads, that much less):
 *   sample_cols -> [all-reduce est] -> sample_fine -> [all-reduce cnt (int64)] -> sample_dist -> ...
 * Worth it from ~8e6 rows per rank (the engine uses s = 5 there). */
int abcdls_fused_sample_cols(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int fine,
                             void* ws, size_t ws_bytes, double** est_out, size_t* est_count, void* stream);
int abcdls_fused_sample_fine(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int fine,
                             void* ws, size_t ws_bytes, int64_t** cnt_out, size_t* cnt_count, void* stream);
int abcdls_fused_sample_dist(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, const double* target,
                             int64_t nacc, int64_t n_global, int world, int fine, void* ws, size_t ws_bytes,
                             double** est_out, size_t* est_count, void* stream);
/* the pass also EMITS every marked row (its d statistics + local row number) into the caller's block
 * cand_ss[ecap][abcdls_fused_emit_pitch(d)] (row-major) / cand_lrow[ecap] (unordered, 64-slot chunks per warp, -1 = unused slot);
 * ecap = abcdls_fused_emit_cap(cap of the candidate list) */
int64_t abcdls_fused_emit_cap(int64_t cap);
int abcdls_fused_emit_pitch(int d);   /* doubles per emitted record = row pitch of cand_ss: 16 (d <= 16) or 32 (d <= 32) */
int abcdls_fused_pass(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int fine,
                      double* cand_ss, int32_t* cand_lrow, int64_t ecap, void* ws, size_t ws_bytes, int* status_dev,
                      int64_t** counts_out, size_t* counts_count, void* stream);
int abcdls_fused_refine(abcdls_handle_t h, const double* ss, int64_t n, int64_t n_global, int d, int64_t ld, int phase,
                        int world, const double* gath_small, const int32_t* gath_cnt, void* ws, size_t ws_bytes,
                        double* med, double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1,
                        size_t* x1_count, void* stream);
/* stage 5 in two halves: abcdls_fused_rows needs only the pass (ascending global candidate rows + their count), so
 * work that only needs the row numbers (abcdls_gather_rows of the parameters on a second stream) can overlap the
 * refinement; abcdls_fused_exact needs the exact MADs.  abcdls_fused_candidates = both. */
int abcdls_fused_rows(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, int64_t row_offset,
                      int64_t cap, int64_t* cand_row, int64_t* n_cand, void* ws, size_t ws_bytes, int* status_dev,
                      void* stream);
int abcdls_fused_exact(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world, const double* mad,
                       const double* target, int64_t cap, double* cand_dist, const double* cand_ss,
                       const int32_t* cand_lrow, int64_t ecap, int64_t* perm, void* ws, size_t ws_bytes, int* status_dev,
                       void* stream);
/* cand_row: ascending global rows; cand_dist: R's exact distance, computed from the emitted block; perm[i] = slot of
 * candidate i in that block; n_cand: device int64 */
int abcdls_fused_candidates(abcdls_handle_t h, const double* ss, int64_t n, int d, int64_t ld, int world,
                            const double* mad, const double* target, int64_t row_offset, int64_t cap, int64_t* cand_row,
                            double* cand_dist, const double* cand_ss, const int32_t* cand_lrow, int64_t ecap,
                            int64_t* perm, int64_t* n_cand, void* ws, size_t ws_bytes, int* status_dev, void* stream);
/* exact ds among the candidates of all ranks (histogram -> bin -> gather -> sort; bins uniform in dist^(2^k)):
 * phase 0 -> all-reduce(sum) x0 (int64); phase 1 -> all-gather x0 (doubles: 8192 values + their count);
 * phase 2 (gath = gathered blocks, NULL on one rank) -> ds_out */
int abcdls_fused_ds(abcdls_handle_t h, int64_t n, int d, int phase, int world, const double* gath,
                    const double* cand_dist, const int64_t* n_cand, int64_t cap, int64_t nacc, double* ds_out, void* ws,
                    size_t ws_bytes, int* status_dev, void** x0, size_t* x0_count, void* stream);
int abcdls_fused_check(abcdls_handle_t h, int64_t n, int d, const double* mad, const double* ds, void* ws,
                       size_t ws_bytes, int* status_dev, void* stream);

/* ---- kernel 3: local-linear adjustment -----------------------------------------------------
 * R: weights (epanechnikov) -> lsfit -> residual recentring -> hcorr -> adj.values.
 * Design is centred on the scaled target; normal equations are accumulated in fp64 with a fixed
 * reduction tree.  M = d + 1.
 *   gather      : accepted rows -> xs (scaled & centred on the target), y, ss, w, dist   [ld = cap]
 *   normal      : partial[ M*M + M*P ] = [X'WX | X'WY] (row-major)      -> all-reduce(sum) across ranks
 *   solve       : beta (M x P, row-major; row 0 = intercept = prediction at the target)
 *   residual    : res = y - X beta
 *   col_stats   : per-column {min,max,sum,sum w x,sum w x^2,sum w}    -> sums feed the_m (all-reduce)
 *   recentre_log: res -= the_m ; logres2 = log(res^2)
 *   normal(with_gram=0, rhs=logres2) -> all-reduce -> solve -> gamma
 *   adjust      : res <- pred.sd*res/pred.si (hcorr) ; adj = beta[0,] + the_m + res
 * n_acc is a device int64 (number of valid rows, <= cap).
 */
int abcdls_gather(abcdls_handle_t h, const double* ss, int64_t ld_ss, const double* param,
                  int64_t ld_param, const double* dist_acc /* distances of the accepted rows */,
                  const int64_t* idx, const int64_t* n_acc, int64_t row_offset, int64_t cap, int d, int P,
                  const double* mad, const double* target, const double* ds, double* xs, double* y,
                  double* ss_out, double* w, void* stream);
/* same outputs; the statistics come from a compact candidate block (cand_ss, leading dimension ld_cand) through
 * perm[slot[r]] (slot: abcdls_accept slot_out; perm: abcdls_fused_candidates, may be NULL = identity; the block is ROW-major, ld_cand doubles per candidate), the
 * parameters from the table through idx[r].
 * One thread per (row, column): d + P independent scattered loads in flight per accepted row. */
/* out[p][i] = table[p][rows[i] - row_offset], i < *n_rows (device int64), leading dimension of out = cap */
int abcdls_gather_rows(abcdls_handle_t h, const double* table, int64_t ld, int P, const int64_t* rows,
                       const int64_t* n_rows, int64_t row_offset, int64_t cap, double* out, void* stream);
/* cand_par (may be NULL): parameters of ALL candidates pre-gathered with abcdls_gather_rows (leading dimension
 * ld_cpar), indexed by slot[r]; when given, the parameter table is not touched */
int abcdls_gather_cand(abcdls_handle_t h, const double* cand_ss, int64_t ld_cand, const int64_t* slot,
                       const int64_t* perm, const double* param, int64_t ld_param, const double* cand_par,
                       int64_t ld_cpar, const double* dist_acc, const int64_t* idx,
                       const int64_t* n_acc, int64_t row_offset, int64_t cap, int d, int P, const double* mad,
                       const double* target, const double* ds, double* xs, double* y, double* ss_out, double* w,
                       void* stream);
/* y[p][r] = param[(idx[r] - row_offset) * row_pitch + p] for the first *n_acc accepted rows: the parameter half of
 * a gather for a ROW-major parameter table (numpy / pandas C order, how ABC-DLS's frames lie in memory:
 * params_unscaled, ABC.py:1825-1875), device or pinned host memory.  One 128-byte request per accepted row (P = 16)
 * instead of P scattered sectors of a column-major table.  abcdls_gather / abcdls_gather_cand accept param == NULL then. */
int abcdls_gather_param_rows(abcdls_handle_t h, const double* param, int64_t row_pitch, int P, const int64_t* idx,
                             const int64_t* n_acc, int64_t row_offset, int64_t cap, double* y, void* stream);
size_t abcdls_normal_workspace_bytes(int d, int P);
int abcdls_normal(abcdls_handle_t h, const double* xs, const double* rhs, const double* w,
                  const int64_t* n_acc, int64_t cap, int d, int P, int with_gram, double* partial,
                  void* ws, size_t ws_bytes, void* stream);
int abcdls_solve(abcdls_handle_t h, const double* gram, const double* rhs, int d, int P, double* beta,
                 int* status_dev, void* stream);
int abcdls_residual(abcdls_handle_t h, const double* xs, const double* y, const double* beta,
                    const int64_t* n_acc, int64_t cap, int d, int P, double* res, void* stream);
/* residual fused with the column sums of the next steps: stat[3][P] = {sum res, sum w res, sum w res^2}
 * (the_m = stat[0] / n; sigma2 = (stat[2] - 2 m stat[1] + m^2 sum w) / sum w).  ws: zero it once. */
size_t abcdls_residual_workspace_bytes(int P);
int abcdls_residual_stats(abcdls_handle_t h, const double* xs, const double* y, const double* beta, const double* w,
                          const int64_t* n_acc, int64_t cap, int d, int P, double* res, double* stat, void* ws,
                          size_t ws_bytes, void* stream);
/* the hcorr fit in one step: res (raw residuals) is centred in place (res -= the_m) and
 * partial[M*P] = X'W log(res^2), without materialising log(res^2) */
int abcdls_normal_logres(abcdls_handle_t h, const double* xs, double* res, const double* the_m, const double* w,
                         const int64_t* n_acc, int64_t cap, int d, int P, double* partial, void* ws, size_t ws_bytes,
                         void* stream);
int abcdls_recentre_log(abcdls_handle_t h, double* res, const double* the_m, const int64_t* n_acc,
                        int64_t cap, int P, double* logres2 /* may be NULL */, void* stream);
int abcdls_adjust(abcdls_handle_t h, const double* xs, double* res, const double* beta,
                  const double* the_m, const double* gamma, const int64_t* n_acc, int64_t cap, int d,
                  int P, int hcorr, double* adj, void* stream);
/* out[c*6 + {0:min, 1:max, 2:sum, 3:sum w*x, 4:sum w*x*x, 5:sum w}] over the first n_acc rows of each
 * column (ld = cap); w may be NULL (=1).  Serves summary() min/max/weighted mean (ABC.py:2833,2837),
 * the_m, sigma2 and the zero-variance-in-region check. */
size_t abcdls_col_stats_workspace_bytes(int ncols);
/* ws: zero it once after allocation; every call leaves it zeroed again */
int abcdls_col_stats(abcdls_handle_t h, const double* values, const double* w, const int64_t* n_acc,
                     int64_t cap, int ncols, double* out, void* ws, size_t ws_bytes, void* stream);

/* P-sized glue of the adjustment (replaces ~35 one-block framework kernels per call).
 * resid_moments: rs[P][6] (col_stats of the residuals) -> sums[3P + 1] = {sum res | sum w res | sum w res^2 | n_acc},
 *   the message of the cross-rank sum.  resid_finish: the_m = colMeans(residuals) (unweighted, R abc()), sigma2 =
 *   sum w (res - the_m)^2 / sum w (gram[0] = sum w), logsig = sum log sigma2 (aic / bic). */
int abcdls_resid_moments(abcdls_handle_t h, const double* rs, const int64_t* n_acc, int P, double* sums, void* stream);
int abcdls_resid_finish(abcdls_handle_t h, const double* sums, const double* gram, int P, double* the_m, double* sig,
                        double* logsig, void* stream);
/* End of a call.  final_pack (multi-rank only): pack[2d + 2P + 8] = {region min | -region max | posterior min |
 * -posterior max | -status bits} for ONE min-exchange, sm[P][4] = stats[:, 2:6] for one sum-exchange.
 * final_small: every small result in one buffer of abcdls_final_small_doubles(d, P, loclinear) doubles:
 *   host[8] = {status, n_acc, ds, region constant (R cond2), every MAD zero, logsig, 0, 0} | mad[d] | stats[P][6] |
 *   and for loclinear (beta != NULL): sigma2[P] | pred[P] = intercept + the_m | coef[(d+1)][P] | coef_h[(d+1)][P] in R's
 *   [1, scaled ss] convention (the fits use a design centred on the scaled target).  pack / sm: the exchanged vectors
 *   (both NULL on one rank: reg / stats / status_dev are read directly). */
size_t abcdls_final_small_doubles(int d, int P, int loclinear);
int abcdls_final_pack(abcdls_handle_t h, const int* status_dev, const double* reg, const double* stats, int d, int P,
                      double* pack, double* sm, void* stream);
int abcdls_final_small(abcdls_handle_t h, const int* status_dev, const int64_t* n_acc, const double* ds,
                       const double* reg, const double* stats, const double* mad, const double* target,
                       const double* logsig /* may be NULL */, const double* beta /* NULL: rejection */,
                       const double* gamma /* NULL: no hcorr */, const double* the_m, const double* sig,
                       const double* pack, const double* sm, int d, int P, double* small_out, void* stream);


/* ---- kernel 3, fused (csrc/tail.cu): the whole tail of abc() as THREE launches (two for rejection) ---------------
 * Each launch ends in its own fixed-order reduction, its cross-rank exchange over the NVLink mailboxes of `peer`
 * (NULL on one rank) and its small solve, all inside the kernel's last thread block:
 *   gather_fit   : accepted rows -> xs | y | ss_out | w (column-major, leading dimension cap; xs / w only for loclinear)
 *                  + X'WX | X'WY + unweighted column sums + region min / max  -> [all-reduce] -> Cholesky -> beta
 *   resid_fit    : (hcorr only) residuals (not materialised) -> X'W log(res^2) -> [all-reduce] -> gamma
 *                  (R's the_m = colMeans(residuals) comes algebraically from the column sums of gather_fit)
 *   adjust_final : adj, res, summary statistics, sigma2 -> [ONE min|sum exchange] -> small_out (layout of
 *                  abcdls_final_small)
 * Statistics source: the candidate block of the one-pass path (cand_ss + slot [+ perm]) or the table (ss, via idx);
 * parameters: column-major (param_cm, ld_param) or row-major (param_rm, row_pitch; device or pinned host) - exactly one.
 * Replaces abcdls_gather* .. abcdls_final_small (16 launches + 3 exchanges) for d, P <= 32.
 * ws: abcdls_tail_workspace_bytes(h, d, P) bytes, zeroed ONCE by the caller (the kernels leave their counters zeroed). */
int abcdls_tail_supported(int d, int P);
size_t abcdls_tail_workspace_bytes(abcdls_handle_t h, int d, int P);
int abcdls_tail_gather_fit(abcdls_handle_t h, abcdls_peer_t peer, const double* cand_ss, int64_t ld_cand,
                           const int64_t* slot, const int64_t* perm, const double* ss, int64_t ld_ss,
                           const double* param_cm, int64_t ld_param, const double* param_rm, int64_t row_pitch,
                           const double* dist_acc, const int64_t* idx, const int64_t* n_acc, int64_t row_offset,
                           int64_t cap, int d, int P, const double* mad, const double* target, const double* ds,
                           int loclinear, double* xs, double* y, double* ss_out, double* w, void* ws, size_t ws_bytes,
                           int* status_dev, void* stream);
int abcdls_tail_resid_fit(abcdls_handle_t h, abcdls_peer_t peer, const double* xs, const double* y, const double* w,
                          const int64_t* n_acc, int64_t cap, int d, int P, void* ws, size_t ws_bytes, int* status_dev,
                          void* stream);
int abcdls_tail_adjust_final(abcdls_handle_t h, abcdls_peer_t peer, const double* xs, const double* y, const double* w,
                             const int64_t* n_acc, const double* ds, int64_t cap, int d, int P, int loclinear,
                             int hcorr, const double* mad, const double* target, double* adj, double* res, void* ws,
                             size_t ws_bytes, double* small_out, int* status_dev, void* stream);

/* ---- per-column mode as ONE pass (csrc/columns.cuh): P independent 1-D abc(method = "rejection") problems sharing the
 * N rows of an N x P statistics table - ABC-DLS's SMC round (ABC.py:2805-2840) and abc_params (1949-1953).
 * For one statistic the distance is monotone in |x - t_j| on either side of the target, so the tolerance cut is a third
 * band of the bracket pass (median band, MAD band, T band); exact or ABCDLS_S_FASTPATH_MISS (then: P separate calls).
 *   sample -> [all-reduce est] -> pass -> [all-reduce counts, tcount] -> refine 0..3 (exchanges as
 *   abcdls_mad_fast_refine) -> exact -> [abcdls_select_* over (P, cap) with cnt, rank nacc - 1 -> sel] -> count ->
 *   [all-gather cnt2 -> quota] -> rowkeys -> [abcdls_select_* per rank, rank quota - 1 -> thr] -> accept (its last
 *   kernel sums / mins the per-column results over the ranks of `peer`).  R's cap rule keeps the FIRST nacc rows of
 *   {dist <= ds} in row order: the members are unordered, so the cut is a row number found by a second selection.
 * nacc_local_hint (this rank's expected share of nacc) sizes the member lists; pass the same value to every call. */
int abcdls_columns_supported(const double* ss, int64_t n, int P, int64_t ld);
size_t abcdls_columns_workspace_bytes(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint);
int abcdls_columns_sample(abcdls_handle_t h, const double* ss, int64_t n, int P, int64_t ld, const double* target,
                          int64_t nacc, int64_t n_global, int64_t nacc_local_hint, int world, void* ws, size_t ws_bytes,
                          double** est_out, size_t* est_count, void* stream);
int abcdls_columns_pass(abcdls_handle_t h, const double* ss, int64_t n, int P, int64_t ld, const double* target,
                        int64_t nacc_local_hint, int world, void* ws, size_t ws_bytes, int64_t** counts_out,
                        size_t* counts_count, int64_t** tcount_out, void* stream);
int abcdls_columns_refine(abcdls_handle_t h, int64_t n, int64_t n_global, int P, int64_t nacc_local_hint, int phase,
                          int world, const double* gath_small, const int32_t* gath_cnt, void* ws, size_t ws_bytes,
                          double* med, double* mad, int* status_dev, void** x0, size_t* x0_count, void** x1,
                          size_t* x1_count, void* stream);
int abcdls_columns_exact(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint, const double* mad,
                         const double* target, void* ws, size_t ws_bytes, double** dist_out, int64_t* cap_out,
                         int64_t* cnt_out, int* status_dev, void* stream);
int abcdls_columns_count(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint, int64_t nacc, const double* sel,
                         const double* mad, const double* target, const int64_t* cnt, void* ws, size_t ws_bytes,
                         int64_t** cnt2_out, int* status_dev, void* stream);
int abcdls_columns_rowkeys(abcdls_handle_t h, int64_t n, int P, int64_t nacc_local_hint, const double* sel,
                           const int64_t* cnt, void* ws, size_t ws_bytes, double** keys_out, void* stream);
int abcdls_columns_accept(abcdls_handle_t h, abcdls_peer_t peer, int64_t n, int P, int64_t nacc_local_hint,
                          const double* sel, const double* med, const double* mad, const int64_t* cnt,
                          const double* thr, const double* par, int64_t par_col_stride, int64_t par_row_stride,
                          int64_t acc_cap, int32_t* acc_row, double* acc_val, double* acc_dist, int64_t* acc_cnt,
                          void* ws, size_t ws_bytes, double* small_out, int* status_dev, void* stream);

/* ---- summary.abc rows that need more than a reduction (csrc/summary.cu; SURVEY 8f-2, printed by ABC.py:844-861) -------
 * wselect: out[p * ntau + t] = weighted order statistic of column p at taus_dev[t] - the smallest x whose cumulated
 *   weight (ascending x) reaches tau * sum(w): quantreg::rq(x ~ 1, tau, weights = w) of summary.abc.  MSB radix select
 *   with fixed-point weight histograms: no sort, deterministic.  x column-major (ld), n rows, w >= 0.
 * col_moments: out[2p] = mean, out[2p + 1] = sum (x - mean)^2 (bw.nrd0's sd).
 * density_mode: the Mode row - argmax over R's density(x, bw, weights = w / wnorm, from, to, n = 512): linear binning,
 *   gaussian convolution (direct, = R's FFT circular convolution), approx() onto [from, to].  lo / up = from - 4 bw /
 *   to + 4 bw.  w may be NULL (equal weights). */
size_t abcdls_wselect_workspace_bytes(int njobs);
int abcdls_wselect(abcdls_handle_t h, const double* x, int64_t ld, const double* w, int64_t n, int P,
                   const double* taus_dev, int ntau, double* out, void* ws, size_t ws_bytes, void* stream);
int abcdls_col_moments(abcdls_handle_t h, const double* x, int64_t ld, int64_t n, int P, double* out, void* stream);
size_t abcdls_density_workspace_bytes(int P);
int abcdls_density_mode(abcdls_handle_t h, const double* x, int64_t ld, const double* w, int64_t n, int P,
                        const double* lo, const double* up, const double* frm, const double* to, const double* bw,
                        const double* wnorm, double* mode_out, void* ws, size_t ws_bytes, void* stream);

/* ---- small tables: one thread block per abc() problem with d = P = 1 (csrc/small.cu) -----------------------------
 * ABC-DLS's per-column mode (ABC.py:1949-1953, 2805-2809) and cv4abc's leave-one-out loop over it (ABC.py:1900-1902)
 * are B independent 1-D problems on test-set sized tables (10^4 rows): launch latency, not bandwidth.  Problem b:
 *   abc(target = hold ? ss[col[b]][hold[b]] : target[b], param = par[col[b]] without row hold[b],
 *       sumstat = ss[col[b]] without row hold[b], tol -> nacc, method = loclinear ? "loclinear" : "rejection")
 * runs entirely in one CTA's shared memory; a whole cross-validation is ONE launch.
 * ss: column-major (ld_ss); par: element (row i, column c) at par[c * par_col_stride + i * par_row_stride] (either
 * layout); n rows BEFORE the hold-out; hold may be NULL (no row removed; then target[B] is read).
 * out[b][8] = {median, mean, ds, mad, min, max, n_acc, status (ABCDLS_S_* bits)} of the posterior sample
 * (summary.abc rows 3, 4, 1, 7: weighted median / mean of adj.values for loclinear, type-7 median / mean of
 * unadj.values for rejection). */
int abcdls_small_supported(int64_t n, int64_t nacc);
int abcdls_small_columns(abcdls_handle_t h, const double* ss, int64_t ld_ss, const double* par, int64_t par_col_stride,
                         int64_t par_row_stride, int64_t n, const int32_t* col, const int64_t* hold,
                         const double* target, int B, int64_t nacc, int loclinear, int hcorr, double* out,
                         void* stream);

/* ---- postpr (rejection): accepted simulations per model ------------------------ */
int abcdls_postpr_counts(abcdls_handle_t h, const int32_t* model_id, const int64_t* idx,
                         const int64_t* n_acc, int64_t row_offset, int64_t cap, int K, int64_t* counts,
                         void* stream);

/* ---- call level (csrc/single.cu): abc(target, param, sumstat, tol, method) on ONE GPU as ONE call ---------------------
 * What the reference binds through rpy2 - abc.abc(target=, param=, sumstat=, tol=, method=) at
 * /root/reference/src/Classes/ABC.py:1950-1953, 1964, 2806-2809 - for hosts that do not want to mirror the stage
 * sequencing of engine.py / stages.py.  Sequences the stage-level entry points of the one-pass path (world = 1) on
 * `stream`; nothing synchronises.  The table must satisfy abcdls_fused_supported (else ABCDLS_E_INVALID: use the stages).
 *   nacc = abcdls_abc_single_nacc(n, tol) = ceiling(n tol); cap = min(nacc, n) is the leading dimension of the outputs:
 *   idx_out[cap] ascending accepted rows, dist_out[cap], y[P][cap] accepted parameters (unadj.values), ss_out[d][cap],
 *   and for loclinear xs[d][cap] (scaled, centred statistics), w[cap], adj[P][cap] (adj.values), res[P][cap];
 *   small_out[abcdls_final_small_doubles(d, P, loclinear)]: {status, n_acc, ds, ...} | mad | stats | sigma2 | pred | coef.
 * param: exactly one of param_cm (column-major P x n, leading dimension ld_param) / param_rm (row-major n x P records,
 * row_pitch doubles apart; may be pinned host memory).  Data-dependent outcomes come back as status bits in small_out[0];
 * ABCDLS_S_FASTPATH_MISS means "not proven complete on this path: rerun through abcdls_mad_fast_* / abcdls_select_*". */
size_t abcdls_abc_single_workspace_bytes(abcdls_handle_t h, int64_t n, int d, int P, double tol);
int64_t abcdls_abc_single_nacc(int64_t n, double tol);
int abcdls_abc_single(abcdls_handle_t h, const double* target, const double* param_cm, int64_t ld_param,
                      const double* param_rm, int64_t row_pitch, const double* sumstat, int64_t ld, int64_t n, int d,
                      int P, double tol, int loclinear, int hcorr, int64_t* idx_out, double* dist_out, double* y,
                      double* ss_out, double* xs, double* w, double* adj, double* res, double* small_out, void* ws,
                      size_t ws_bytes, void* stream);

/* Results out of the buffers of a replayed CUDA graph in ONE launch: up to 4 segments (rows[i] x cols[i] 8-byte
 * elements, source pitch ld[i]) are packed densely, one after the other, into dst. */
int abcdls_copy_out(abcdls_handle_t h, int nseg, const void* const* src, const int64_t* ld, const int64_t* rows,
                    const int64_t* cols, void* dst, void* stream);
/* the end of a postpr() call in one block: counts and status of all ranks in ONE exchange (peer; NULL on one rank), then
 * small_out[8 + d + K] = {status, n_acc, ds, every MAD zero, 0, 0, 0, 0 | mad[d] | counts[K] (global, as doubles)} */
int abcdls_postpr_final(abcdls_handle_t h, abcdls_peer_t peer, const int* status_dev, const int64_t* n_acc,
                        const double* ds, const double* mad, int d, const int64_t* counts, int K, double* small_out,
                        void* stream);

/* ---- kernel 4: SMC ------------------------------------------------------------------------
 * oldrange  = per-column min/max over ALL simulated parameter rows   (ABC.py:2714-2716)
 * box filter = rows with every parameter in [lo, hi] inclusive        (ABC.py:2974-2990)
 */
size_t abcdls_smc_workspace_bytes(int64_t n, int P);
int abcdls_smc_oldrange(abcdls_handle_t h, const double* param_all, int64_t n, int P, int64_t ld,
                        double* minmax /* [2*P]: mins then maxes */, void* ws, size_t ws_bytes,
                        void* stream);
int abcdls_smc_box_filter(abcdls_handle_t h, const double* param_all, int64_t n, int P, int64_t ld,
                          const double* lo, const double* hi, int64_t row_offset, int64_t cap,
                          int64_t* idx_out, int64_t* n_out, void* ws, size_t ws_bytes, int* status_dev,
                          void* stream);

/* ---- collectives over NVLink peer memory (csrc/peer.cu) ---------------------------------------
 * The exchanges of the path are tiny and latency bound.  Every rank owns a mailbox (cudaMalloc + CUDA
 * IPC) that all peers map; one kernel per exchange writes the contribution, pushes a sequence number to the peers,
 * waits for theirs and reduces / concatenates the peers' slots in RANK ORDER over NVLink.  No host involvement and
 * no NCCL on the data path: the kernels are ordinary graph nodes.  Bootstrap (exchanging the IPC handles) is the
 * caller's job (the Python host uses torch.distributed for it).
 *   create (per rank) -> all-gather the handles by any channel -> connect -> barrier -> exchanges ... -> destroy */
size_t abcdls_peer_handle_bytes(void);
int abcdls_peer_create(abcdls_peer_t* out, int rank, int world, size_t slot_bytes, void* handle_out);
int abcdls_peer_connect(abcdls_peer_t p, const void* all_handles /* world x handle_bytes, rank order */);
int abcdls_peer_destroy(abcdls_peer_t p);
/* profiling aid: with ABCDLS_PEER_TRACE=1 in the environment at abcdls_peer_create, every exchange records four
 * %globaltimer stamps (entry, own contribution published, every peer's flag seen, exit); out: 4096 x 4 uint64 ns (record
 * k % 4096 = exchange k), n_exchanges: exchanges completed.  Returns the record capacity, 0 when tracing is off. */
int abcdls_peer_trace_read(abcdls_peer_t p, uint64_t* out, uint64_t* n_exchanges);
const char* abcdls_peer_last_error(abcdls_peer_t p);
size_t abcdls_peer_slot_bytes(abcdls_peer_t p);
/* in place; dtype 0 = double, 1 = int64; op 0 = sum, 1 = min, 2 = max; count * 8 <= slot bytes */
int abcdls_peer_allreduce(abcdls_peer_t p, void* data, size_t count, int dtype, int op, void* stream);
/* dst[world * bytes] <- every rank's src[bytes] (multiple of 8, <= slot bytes), rank order */
int abcdls_peer_allgather(abcdls_peer_t p, const void* src, size_t bytes, void* dst, void* stream);
/* nj ragged rows per rank: src_vals[nj][cap] with src_cnt[j] existing entries in row j -> dst_vals[world][nj][cap],
 * dst_cnt[world][nj]; only the existing entries cross NVLink.  (nj + nj * cap) * 8 <= slot bytes */
int abcdls_peer_allgather_ragged(abcdls_peer_t p, const double* src_vals, const uint32_t* src_cnt, int nj, int64_t cap,
                                 double* dst_vals, uint32_t* dst_cnt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ABCDLS_H */
