"""Host-side sequencing of the ABC acceptance-and-adjustment path on B200.

One process per GPU.  Rows are sharded in contiguous rank-major blocks (global row = row_offset +
local row), tables live in HBM column-major (``(cols, n_local)`` torch tensors with unit stride along
rows — R's ``as.matrix`` layout).  Every arithmetic step is a CUDA kernel behind the C ABI
(``stages.py``); this file only orders the stages and puts the small collectives between them:

    column median/MAD  : radix-select passes, histogram all-reduce per pass when world > 1
    distance           : local
    threshold ds       : radix-select on dist (same all-reduce)
    cap rule           : all-gather of one int64 per rank -> exclusive scan -> local quota
    normal equations   : all-reduce of the (d+1) x (d+1+P) fp64 partial
    the_m / sigma2     : all-reduce of P(+1) sums
    summary            : all-reduce min / max / sums

Reference boundary: src/Classes/ABC.py:1950-1953, 1964 (abc.abc), 919 (abc.postpr), 2806-2840 (SMC
min/max), with the arithmetic of CRAN abc 2.2.1.
"""
from __future__ import annotations

import math
import os
import time
import weakref
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _capi
from .comm import Comm


class AbcError(RuntimeError):
    """Mirrors an R ``stop()`` inside abc()/postpr() (surfaced by rpy2 as RRuntimeError)."""


class _Deferred:
    """A result array that still lives in the buffers of a replayed CUDA graph.  It is copied out on first access; if
    the same plan has been replayed in the meantime the data is gone and the access raises instead of returning the
    newer call's values."""

    def __init__(self, plan, view, post=None):
        self.plan, self.gen, self.view, self.post = plan, plan["gen"], view, post

    def resolve(self):
        if self.plan["gen"] != self.gen:
            raise RuntimeError("this array of an earlier abc() result was not read before the next call with the same "
                               "tables reused the graph's buffers; read it first, or set engine.lazy_results = False")
        t = (self.view() if callable(self.view) else self.view).clone()
        return self.post(t) if self.post is not None else t


class _Lazy:
    """A field of a result that is a cheap view of data the result already OWNS (its packed copy-out block): built on
    first access, so that the per-call host time between two graph launches stays short."""

    def __init__(self, fn):
        self.fn = fn

    def resolve(self):
        return self.fn()


@dataclass
class AbcResult:
    method: str
    tol: float
    nacc: int                      # ceiling(N * tol), global
    n_global: int
    ds: float
    mad: torch.Tensor              # (d,)  R mad() of every column (0 => column left unscaled)
    idx: torch.Tensor              # (n_local_acc,) ascending GLOBAL row numbers held by this rank
    unadj_values: torch.Tensor     # (n_local_acc, P)
    ss: torch.Tensor               # (n_local_acc, d)  unscaled accepted summary statistics
    weights: torch.Tensor          # (n_local_acc,)    1 for rejection, Epanechnikov otherwise
    dist_accepted: torch.Tensor    # (n_local_acc,)
    adj_values: Optional[torch.Tensor] = None
    residuals: Optional[torch.Tensor] = None
    coef: Optional[torch.Tensor] = None      # (d+1, P) fit1, R convention (uncentred design [1, scaled ss])
    coef_h: Optional[torch.Tensor] = None    # (d+1, P) fit2, R convention
    pred: Optional[torch.Tensor] = None      # (P,) prediction at the target incl. the_m
    sigma2: Optional[torch.Tensor] = None
    aic: Optional[float] = None
    bic: Optional[float] = None
    stats: Optional[torch.Tensor] = None     # (P, 6) global {min,max,sum,sum wx,sum wx^2,sum w} of the posterior sample
    dist: Optional[torch.Tensor] = None      # (n_local,) full local distance vector when requested
    warnings: List[str] = field(default_factory=list)
    numparam: int = 0
    numstat: int = 0

    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if isinstance(v, (_Deferred, _Lazy)):
            v = v.resolve()
            object.__setattr__(self, name, v)
        return v

    @property
    def values(self) -> torch.Tensor:
        """The posterior sample summary() works on: adj.values if present else unadj.values."""
        return self.adj_values if self.adj_values is not None else self.unadj_values

    def minmax(self):
        """(min, max) per parameter over the global posterior sample — what ABC-DLS's SMC reads as
        summary(res)[0] and summary(res)[6] (ABC.py:2833, 2837)."""
        return self.stats[:, 0].clone(), self.stats[:, 1].clone()


@dataclass
class PostprResult:
    """``ss`` / ``values`` (the accepted rows' statistics and model ids) may be lazy: R's summary.postpr only needs the
    counts, so the rows are gathered from the caller's tables on first access."""
    models: List[str]
    counts: torch.Tensor           # (K,) int64 accepted simulations per model (global)
    pred: torch.Tensor             # (K,) proportions
    nacc: int
    n_global: int
    ds: float
    mad: torch.Tensor
    idx: torch.Tensor
    ss: torch.Tensor
    values: torch.Tensor           # (n_local_acc,) int32 model ids of accepted rows, row order
    dist: Optional[torch.Tensor] = None

    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if isinstance(v, _Lazy):
            v = v.resolve()
            object.__setattr__(self, name, v)
        elif callable(v) and name in ("ss", "values"):
            v = v()
            object.__setattr__(self, name, v)
        return v

    def bayes_factors(self) -> torch.Tensor:
        return self.pred[:, None] / self.pred[None, :]


def _as_table(t: torch.Tensor, row_major_ok: bool = False) -> torch.Tensor:
    """(cols, rows) float64 view with unit stride along rows.  The tensor stays where it is: a PINNED host table is
    legal for `param` (the gather kernel then reads the accepted rows zero-copy over PCIe/UVA).  `param` may also be
    ROW-major - a (P, n) view with strides (1, pitch), i.e. ``array_n_by_P.t()``, how numpy / pandas hand the frame
    over - and is then left as it is: its accepted rows are gathered as contiguous 8 P-byte records."""
    if t.dim() == 1:
        t = t[None, :]
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if t.stride(1) != 1:
        if row_major_ok and t.stride(0) == 1 and t.stride(1) >= t.shape[0]:
            return t
        t = t.contiguous()
    return t


class _SegmentedGraph:
    """The device work of one call as CUDA-graph segments.  Kernels, memsets and torch ops are captured; every
    collective ends the running segment, runs eagerly and is replayed eagerly between two graph launches (NCCL is
    never captured).  A single-rank call is one segment = one graph launch.  All segments allocate from one pool."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.pool = torch.cuda.graph_pool_handle()
        self.items = []
        self.cur = None

    def begin(self):
        g = torch.cuda.CUDAGraph()
        # thread_local: other threads (NCCL watchdog, NVML sampler) may keep calling the CUDA API while we capture
        g.capture_begin(pool=self.pool, capture_error_mode="thread_local")
        self.cur = g

    def end(self):
        self.cur.capture_end()
        self.items.append(self.cur)
        self.cur = None

    def split(self, fn):
        self.end()
        fn()
        self.items.append(fn)
        self.begin()

    def replay(self):
        for it in self.items:
            if isinstance(it, torch.cuda.CUDAGraph):
                it.replay()
            else:
                it()


class AbcEngine:
    """Drop-in compute engine for abc()/postpr() on device tables.  See ``rabc.py`` for the facade that
    keeps rpy2's call surface."""

    def __init__(self, device=None, group=None, stages=None):
        if stages is None:
            from .stages import CudaStages
            if device is None:
                device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
            if device is None:
                raise RuntimeError("abc_dls_b200: no CUDA device — the engine has no CPU path")
            stages = CudaStages(torch.device(device))
        self.k = stages
        self.device = stages.device
        self.comm = Comm(group, self.device)
        self.timers = None   # set to {} to record CUDA-event pairs per named stage (bench.py)
        self.fast_misses = 0  # how often a sampled bracket missed and the next path of the ladder reran the call
        self.path_calls = {"fused": 0, "two_pass": 0, "exact": 0}   # front-half path taken, per call
        self.host_enqueue_ms = []   # per abc() call: host time spent enqueueing before the one blocking read
        self.use_graphs = True      # replay the device work of a repeated call as one CUDA graph
        self.copy_results = True    # results of a replayed graph are copied out of the graph's buffers ...
        self.lazy_results = True    # ... the rarely read ones (ss, residuals, unadj of loclinear, dist) on first access
        self.max_graphs = 32        # per-column mode of ABC-DLS makes P (19, 24) calls with distinct tables in a row
        # Gather the parameters of ALL candidates on a second stream during the refinement.  Measured at 1e8 x 16: the
        # scattered reads (1.6x more rows) slow the memory-bound refinement kernels by more than the later gather
        # saves (4.94 -> 5.2 ms per call), so it stays off.
        self.pregather = False
        self.fused_tail = os.environ.get("ABCDLS_NO_TAIL", "0") != "1"   # csrc/tail.cu instead of the 16-launch tail
        self.use_columns = os.environ.get("ABCDLS_NO_COLUMNS", "0") != "1"   # csrc/columns.cuh for per-column summaries
        self._side = None
        if os.environ.get("ABCDLS_NO_GRAPH", "0") == "1":
            self.use_graphs = False
        self._graphs, self._seen = {}, {}
        self.host_profile = dict(calls=0, enqueue_us=0.0, wait_us=0.0, result_us=0.0, between_calls_us=0.0)
        self._t_last_return = None
        self._last_call = None
        self._last_postpr = None
        if hasattr(stages, "timer"):
            stages.timer = self._t

    class _Timed:
        def __init__(self, eng, name):
            self.eng, self.name = eng, name

        def __enter__(self):
            if self.eng.timers is not None:
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record(torch.cuda.current_stream(self.eng.device))

        def __exit__(self, *exc):
            if self.eng.timers is not None:
                self.b.record(torch.cuda.current_stream(self.eng.device))
                self.eng.timers.setdefault(self.name, []).append((self.a, self.b))
            return False

    def _t(self, name):
        return AbcEngine._Timed(self, name)

    def timer_ms(self):
        """name -> list of elapsed ms for every recorded (start, end) pair; call after a synchronize."""
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in (self.timers or {}).items()}

    # ---- small helpers ----------------------------------------------------------------------------
    def _shard_info(self, n_local: int, flag: int = 1):
        """(global rows, first global row of this rank, rows per rank, AND of `flag` over ranks)."""
        counts = self.comm.allgather_ints([n_local, flag])  # list per rank
        ns = [c[0] for c in counts]
        return sum(ns), sum(ns[: self.comm.rank]), ns, all(c[1] for c in counts)

    def _select(self, table: torch.Tensor, n_local: int, center, k0, k1, counts=None, local: bool = False) -> torch.Tensor:
        """Exact global order statistics k0,k1 (0-based) of every row of ``table`` (njobs, n_local); counts (device
        int64 per job, optional): only the first counts[j] entries of row j exist.  local: order statistics of THIS
        rank's entries (no histogram exchange)."""
        k = self.k
        njobs, ld = table.shape[0], table.stride(0)
        ws = k.select_workspace(njobs)
        k.select_begin(ws, njobs, table, n_local, ld, counts, center, k0, k1)
        hist = k.select_hist_tensor(ws, njobs) if (self.comm.world > 1 and not local) else None
        for p in range(_capi.SELECT_PASSES):
            with self._t(f"select_hist[{njobs}x{n_local}]"):
                k.select_hist(ws, njobs, n_local, p)
            if hist is not None:
                self.comm.allreduce_sum(hist)
            k.select_pick(ws, njobs, p)
        out = k.new(njobs, 2)
        k.select_end(ws, njobs, out)
        return out

    def column_mad(self, ss: torch.Tensor, n_global: int):
        """R median() and mad() of every column, exactly (even N: mean of the two middle values)."""
        k = self.k
        d, n_local = ss.shape
        k0, k1 = (n_global - 1) // 2, n_global // 2
        med, mad = k.new(d), k.new(d)
        r = self._select(ss, n_local, None, k0, k1)
        k.median_from_ranks(r, d, 1.0, med)
        r = self._select(ss, n_local, med, k0, k1)
        k.median_from_ranks(r, d, 1.4826, mad)
        return med, mad

    # ---- shared front half: scaling, distance, tolerance cut ----------------------------------
    FAST_MIN_ROWS = 1 << 18   # rows per rank; below this the 6-pass radix select (launch bound, always exact) runs
    use_fused = os.environ.get("ABCDLS_NO_FUSED", "0") != "1"   # one-pass front half (csrc/fused.cuh) when the table qualifies
    FINE_SHIFT = int(os.environ.get("ABCDLS_FINE", "5"))               # second-stage sample: 1 line in 2^5 (0 = off)
    FINE_MIN_ROWS = int(os.environ.get("ABCDLS_FINE_MIN", str(1 << 23)))

    def _plan(self, ss, tol, keep_dist=False, level=0, shard=None):
        """Host-side decisions for one call (the only place that may need a host collective: shard sizes).
        level 0: one-pass front half if the table qualifies; 1: two-pass sampled brackets; 2: 6-pass radix select.
        shard = (n_global, row_offset): the caller vouches for the standard contiguous split (rank r holds rows
        [r * ceil(N / world), ...)) with the same table layout on every rank, which saves the blocking exchange."""
        k = self.k
        d, n_local = ss.shape
        can_fuse = (self.use_fused and level == 0 and not keep_dist and hasattr(k, "fused_supported")
                    and k.fused_supported(ss, n_local, d, ss.stride(0)))
        if shard is not None and self.comm.world > 1:
            n_global, row_offset = int(shard[0]), int(shard[1])
            per = -(-n_global // self.comm.world)
            ns = [max(0, min(per, n_global - r * per)) for r in range(self.comm.world)]
            if ns[self.comm.rank] != n_local or row_offset != self.comm.rank * per:
                raise AbcError("shard=(n_global, row_offset) does not describe the standard contiguous split")
            all_fuse = can_fuse
            if (self.use_fused and level == 0 and not keep_dist and min(ns) >= 16384 and d <= 32 and not can_fuse):
                raise AbcError("shard= promises the same 16-byte aligned, even-pitch table layout on every rank")
        else:
            n_global, row_offset, ns, all_fuse = self._shard_info(n_local, int(can_fuse))
        nacc = int(math.ceil(n_global * tol))
        if not (1 <= nacc <= n_global):
            raise AbcError("'tol' must give 1 <= ceiling(N * tol) <= N")
        fast = level < 2 and min(ns) >= self.FAST_MIN_ROWS   # every rank must take the same path
        fused = fast and all_fuse
        # second-stage counting sample of the one-pass path (csrc/fused.cuh): from FINE_MIN_ROWS rows on EVERY rank
        fine_min = self.FINE_MIN_ROWS * (2 if self.comm.world > 1 else 1)   # several ranks: one more exchange to pay for
        fine = self.FINE_SHIFT if (fused and min(ns) >= fine_min) else 0
        return dict(n_global=n_global, row_offset=row_offset, nacc=nacc, fast=fast, fused=fused, keep_dist=keep_dist, fine=fine,
                    path="fused" if fused else ("two_pass" if fast else "exact"))

    def _front(self, target, ss, status, pl, par=None):
        """MAD scaling -> distance -> exact threshold ds -> R's cap rule (device work only).  Returns the accepted rows
        (ascending global row numbers, their distances) of this rank.

        Paths (csrc/fused.cuh, csrc/bracket.cu, csrc/select.cu): one-pass front half (ONE table read), two-pass
        sampled brackets, 6-pass radix select.  The first two verify themselves on the device and raise
        S_FASTPATH_MISS; abc() then reruns one level down the ladder."""
        k = self.k
        d, n_local = ss.shape
        n_global, row_offset, nacc = pl["n_global"], pl["row_offset"], pl["nacc"]
        fast, fused, keep_dist = pl["fast"], pl["fused"], pl["keep_dist"]
        cap = min(nacc, n_local)
        idx = k.new(cap, dtype=torch.int64)
        dacc = k.new(cap)
        state = torch.zeros(4, dtype=torch.int64, device=self.device)      # one fill for the scalar outputs
        n_acc, le, n_cand0 = state[0:1], state[1:2], state[2:3]
        quota = torch.full((1,), nacc, dtype=torch.int64, device=self.device)
        comm = self.comm if self.comm.world > 1 else None
        cand_ss = slot = perm = cand_par = None
        if fused:
            # ONE table read: medians + MADs and the candidate rows of the cut; exact distances only for the candidates
            med, mad = k.new(d), k.new(d)
            ccap = min(n_local, 2 * nacc + 65536)
            dist = None
            cand_row, cand_dist = k.new(ccap, dtype=torch.int64), k.new(ccap)
            ecap = k.fused_emit_cap(ccap)                     # the pass emits candidate rows in 64-slot chunks per warp
            cand_ss = k.new(ecap, k.fused_emit_pitch(d))      # row-major: 16 (d <= 16) or 32 doubles per emitted row
            cand_lrow = k.new(ecap, dtype=torch.int32)
            perm = k.new(ccap, dtype=torch.int64)
            slot = k.new(cap, dtype=torch.int64)
            n_cand = n_cand0
            pre = None
            if (par is not None and par.device.type == "cuda" and self.pregather and par.stride(1) == 1
                    and (self.comm.world == 1 or self.comm.peer is not None)):   # NCCL segments cannot span a fork
                # device-resident parameters: gather them for ALL candidates on a second stream while this one refines
                if self._side is None:
                    self._side = torch.cuda.Stream(self.device)
                cand_par = k.new(par.shape[0], ccap)
                pre = (par, cand_par, self._side)
            with self._t("fused_front"):
                k.fused_front(ss, n_local, n_global, d, ss.stride(0), target, nacc, row_offset, ccap, med, mad, cand_row,
                              cand_dist, cand_ss, cand_lrow, perm, n_cand, status, comm, pregather=pre,
                              fine=pl.get("fine", 0))
            with self._t("fused_ds"):
                ds = k.new(1)
                k.fused_ds(n_local, d, nacc, ccap, cand_dist, n_cand, ds, status, comm)
                k.fused_check(n_local, d, mad, ds, status)
            with self._t("accept"):
                ws = k.accept_workspace(ccap)
                k.count_le(cand_dist, ccap, n_cand, ds, le, ws)
                if comm is not None:
                    below = comm.allgather(le).view(-1)[: comm.rank].sum()
                    quota = (nacc - below).clamp(min=0).view(1)
                k.accept(cand_dist, ccap, n_cand, cand_row, ds, quota, row_offset, cap, idx, dacc, n_acc, ws, status,
                         slot_out=slot)
        elif fast:
            med, mad = k.new(d), k.new(d)
            with self._t("mad_fast"):
                k.mad_fast(ss, n_local, n_global, d, ss.stride(0), med, mad, status, comm)
            ccap = min(n_local, 2 * nacc + 65536)
            dist = k.new(n_local) if keep_dist else None
            cand_row, cand_dist = k.new(ccap, dtype=torch.int64), k.new(ccap)
            n_cand = n_cand0
            ds = k.new(1)
            with self._t("dist_fast"):
                k.dist_fast(ss, n_local, n_global, d, ss.stride(0), mad, target, nacc, row_offset, ccap, dist,
                            cand_row, cand_dist, n_cand, ds, status, comm)
            with self._t("accept"):
                ws = k.accept_workspace(ccap)
                k.count_le(cand_dist, ccap, n_cand, ds, le, ws)
                if comm is not None:   # R's cap rule across ranks: lower ranks take their rows first
                    below = comm.allgather(le).view(-1)[: comm.rank].sum()
                    quota = (nacc - below).clamp(min=0).view(1)
                k.accept(cand_dist, ccap, n_cand, cand_row, ds, quota, row_offset, cap, idx, dacc, n_acc, ws, status)
        else:
            _, mad = self.column_mad(ss, n_global)
            dist = k.new(n_local)
            with self._t("scale_dist"):
                k.scale_dist(ss, n_local, d, ss.stride(0), mad, target, dist, status)
            ds = self._select(dist[None, :], n_local, None, nacc - 1, nacc - 1)[0, :1]  # (1,) device
            ws = k.accept_workspace(n_local)
            k.count_le(dist, n_local, None, ds, le, ws)
            if self.comm.world > 1:
                allc = self.comm.allgather(le).view(-1)
                below = allc[: self.comm.rank].sum()
                quota = (nacc - below).clamp(min=0).view(1)
            k.accept(dist, n_local, None, None, ds, quota, row_offset, cap, idx, dacc, n_acc, ws, status)
        return dict(row_offset=row_offset, mad=mad, dist=dist, ds=ds, cap=cap, idx=idx, n_acc=n_acc, dacc=dacc, fast=fast,
                    fused=fused, cand_ss=cand_ss, slot=slot, perm=perm, cand_par=cand_par)

    def _zero_variance(self, ss) -> bool:
        """R: stop if every column of sumstat has a single unique value."""
        k = self.k
        d, n_local = ss.shape
        mm = k.new(2 * d)
        k.smc_oldrange(ss, n_local, d, ss.stride(0), mm)
        mn, mx = mm[:d].clone(), mm[d:].clone()
        self.comm.allreduce_min(mn)
        self.comm.allreduce_max(mx)
        return bool((mn == mx).all().item())

    # ---- abc() ----------------------------------------------------------------------------------
    def abc(self, target: torch.Tensor, param: torch.Tensor, sumstat: torch.Tensor, tol: float, method: str,
            hcorr: bool = True, kernel: str = "epanechnikov", keep_dist: bool = False,
            shard=None, _level: int = 0) -> AbcResult:
        """``abc::abc`` for method in {rejection, loclinear} on this rank's shard.

        target (d,), param (P, n_local), sumstat (d, n_local): float64 CUDA tensors, column-major tables.
        `param` may instead be a PINNED host tensor: only the accepted rows (tol * N) are ever read, so the table is
        left in host memory and gathered zero-copy (saves the H2D copy of N x P doubles).
        """
        _t0 = time.perf_counter()
        if method not in ("rejection", "loclinear"):
            raise AbcError("Method must be rejection or loclinear (neuralnet / ridge stay with R's nnet)")
        if kernel != "epanechnikov":
            raise AbcError("only kernel='epanechnikov' (the R default ABC-DLS uses) is implemented")
        k = self.k
        if not (target.dtype == torch.float64 and target.device == self.device and target.dim() == 1
                and target.is_contiguous()):
            target = target.to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
        # A repeated call on the SAME table objects (the per-column loops, cross-validation, SMC rounds of ABC-DLS call
        # in a row) reuses the validated views, the plan and the graph key of the previous one: the host time before the
        # graph launch is GPU idle time.  Only weak references are kept (the engine must not pin the caller's tables): a
        # live weak reference to the very same objects + equal addresses / shapes / strides means the same tables.
        sig = (id(sumstat), id(param), sumstat.data_ptr(), param.data_ptr(), tuple(sumstat.shape), tuple(param.shape),
               sumstat.stride(), param.stride(), float(tol), method, bool(hcorr), bool(keep_dist), int(_level),
               None if shard is None else (int(shard[0]), int(shard[1])))
        # Several ranks without `shard=`: _plan runs a collective (shard sizes), which every rank must enter or none -
        # a memo hit on one rank only would hang the others, so the memo is used only where _plan is local.
        local_plan = self.comm.world == 1 or shard is not None
        memo = self._last_call if local_plan else None
        if memo is not None and memo[0] == sig and memo[1]() is sumstat and memo[2]() is param:
            ss, par = sumstat, param
            _, _, _, d, n_local, P, pl, key = memo
            if target.numel() != d:
                raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
        else:
            ss, par = _as_table(sumstat), _as_table(param, row_major_ok=True)
            d, n_local = ss.shape
            P = par.shape[0]
            if target.numel() != d:
                raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
            if par.shape[1] != n_local:
                raise AbcError("'param' and 'sumstat' must have the same number of rows")
            if par.device.type == "cpu":
                if self.device.type == "cuda" and not par.is_pinned():
                    par = par.to(self.device)      # pageable host memory cannot be mapped: stage it
            elif par.device != self.device:
                par = par.to(self.device)
            if ss.device != self.device:
                ss = ss.to(self.device)
            pl = self._plan(ss, tol, keep_dist, _level, shard)
            # The device work of one call is a fixed sequence of ~45 launches whose shapes depend only on the key below,
            # so it is captured into a CUDA graph the second time a key is seen and replayed afterwards.
            key = ("abc", ss.data_ptr(), tuple(ss.shape), ss.stride(0), par.data_ptr(), tuple(par.shape), par.stride(0),
                   par.stride(1), par.device.type, float(tol), method, bool(hcorr), bool(keep_dist), int(_level),
                   self.comm.world, pl["n_global"], pl["row_offset"])
            self._last_call = ((sig, weakref.ref(sumstat), weakref.ref(param), d, n_local, P, pl, key)
                               if (ss is sumstat and par is param) else None)
        self.path_calls[pl["path"]] += 1

        o, plan = self._run(key, target, lambda tgt: self._abc_device(tgt, ss, par, pl, method, hcorr),
                            graphable=not keep_dist)
        from_graph = plan is not None

        _t1 = time.perf_counter()
        self.host_enqueue_ms.append((_t1 - _t0) * 1e3)
        if len(self.host_enqueue_ms) > 64:
            del self.host_enqueue_ms[:32]
        host = o["small"][:8].cpu()    # the one blocking read: everything data dependent
        _t2 = time.perf_counter()
        host = host.tolist()           # one conversion: plain floats from here on
        st, n_loc, ds_h = int(host[0]), int(host[1]), float(host[2])    # status is already OR-ed over the ranks
        warnings: List[str] = []
        if (st & _capi.S_FASTPATH_MISS) and pl["fast"]:
            # fallback ladder: one-pass -> two-pass sampled brackets -> 6-pass radix select (always exact)
            self.fast_misses += 1
            return self.abc(target, param, sumstat, tol, method, hcorr, kernel, keep_dist, shard,
                            _level=1 if pl["fused"] else 2)
        if st & _capi.S_NONFINITE:
            raise AbcError("non-finite values in 'sumstat'/'target' (R's NA-row handling is not implemented)")
        if host[4] != 0 and self._zero_variance(ss):
            raise AbcError("Zero variance in the summary statistics.")
        if host[3] != 0:
            msg = "Zero variance in the summary statistics in the selected region."
            if method != "rejection":
                raise AbcError(msg + " Try: checking summary statistics, choosing larger tolerance, or rejection method.")
            warnings.append(msg + " Check summary statistics, consider larger tolerance.")
        if st & _capi.S_PEER_TIMEOUT:
            raise RuntimeError("a rank did not take part in an exchange of this call (NVLink mailbox wait timed out)")
        if st & _capi.S_SINGULAR:
            raise AbcError("lsfit: X matrix deemed to be singular (collinear summary statistics in the accepted region)")
        if st & _capi.S_NONFINITE_FIT:
            raise AbcError("lsfit: NA/NaN/Inf in 'y' (hcorr: a residual is exactly zero, log(residuals^2) = -Inf)")
        if st & _capi.S_CAPACITY:
            raise AbcError("internal: accepted-row buffer overflow")

        # Results of a replayed graph live in the graph's buffers (overwritten by the next call with the same key): what
        # a caller normally reads (rows, posterior sample, weights, summary statistics) is copied out now, the rest on
        # first access (_Deferred).
        nacc = pl["nacc"]
        lin = method == "loclinear"
        M = d + 1
        tr = lambda t: t.t()
        if from_graph and self.copy_results and self.lazy_results and hasattr(k, "copy_out"):
            # ONE launch packs small | idx | weights | posterior sample into a block this result owns; every field is a
            # view of it built on first access (the host time spent here delays the next call's graph launch)
            ns = o["small"].numel()
            main = o["adj"] if lin else o["y"]
            segs = [(o["small"], 1, ns), (o["idx"], 1, n_loc)] + ([(o["w"], 1, n_loc)] if lin else []) + [(main, P, n_loc)]
            flat = k.copy_out(segs, memo=plan)
            o_w = ns + n_loc
            o_main = o_w + (n_loc if lin else 0)
            small = flat[:ns]
            later = lambda t, post=None: _Deferred(plan, t, post)
            sample = _Lazy(lambda: flat[o_main:o_main + P * n_loc].view(P, n_loc).t())
            q = 8 + d + 6 * P
            r = AbcResult(method=method, tol=tol, nacc=nacc, n_global=pl["n_global"], ds=ds_h,
                          mad=_Lazy(lambda: small[8:8 + d]),
                          idx=_Lazy(lambda: flat[ns:ns + n_loc].view(torch.int64)),
                          unadj_values=(later(lambda: o["y"][:, :n_loc], tr) if lin else sample),
                          ss=later(lambda: o["ss_out"][:, :n_loc], tr),
                          weights=(_Lazy(lambda: flat[o_w:o_w + n_loc]) if lin else
                                   _Lazy(lambda: torch.ones(n_loc, dtype=torch.float64, device=self.device))),
                          dist_accepted=later(lambda: o["dacc"][:n_loc]),
                          stats=_Lazy(lambda: small[8 + d:8 + d + 6 * P].view(P, 6)), warnings=warnings, numparam=P,
                          numstat=d, dist=None)
            if lin:
                r.adj_values, r.residuals = sample, later(lambda: o["res"][:, :n_loc], tr)
                r.sigma2, r.pred = _Lazy(lambda: small[q:q + P]), _Lazy(lambda: small[q + P:q + 2 * P])
                r.coef = _Lazy(lambda: small[q + 2 * P:q + 2 * P + M * P].view(M, P))
                if hcorr:
                    r.coef_h = _Lazy(lambda: small[q + 2 * P + M * P:q + 2 * P + 2 * M * P].view(M, P))
        else:
            if from_graph and self.copy_results:
                own = lambda t: t.clone()
                later = (lambda t, post=None: _Deferred(plan, t, post)) if self.lazy_results else \
                        (lambda t, post=None: post(t.clone()) if post else t.clone())
            else:
                own = lambda t: t
                later = lambda t, post=None: post(t) if post else t
            small = own(o["small"])         # mad | stats | sigma2 | pred | coef | coef_h behind the 8-double host word
            mad, stats = small[8:8 + d], small[8 + d:8 + d + 6 * P].view(P, 6)
            r = AbcResult(method=method, tol=tol, nacc=nacc, n_global=pl["n_global"], ds=ds_h, mad=mad,
                          idx=own(o["idx"][:n_loc]),
                          unadj_values=(later(o["y"][:, :n_loc], tr) if lin else own(o["y"][:, :n_loc]).t()),
                          ss=later(o["ss_out"][:, :n_loc], tr),
                          weights=(own(o["w"][:n_loc]) if lin
                                   else torch.ones(n_loc, dtype=torch.float64, device=self.device)),
                          dist_accepted=later(o["dacc"][:n_loc]), stats=stats, warnings=warnings, numparam=P,
                          numstat=d, dist=o["dist"] if keep_dist else None)
            if lin:
                q = 8 + d + 6 * P
                r.adj_values, r.residuals = own(o["adj"][:, :n_loc]).t(), later(o["res"][:, :n_loc], tr)
                r.sigma2, r.pred = small[q:q + P], small[q + P:q + 2 * P]
                r.coef = small[q + 2 * P:q + 2 * P + M * P].view(M, P)
                if hcorr:
                    r.coef_h = small[q + 2 * P + M * P:q + 2 * P + 2 * M * P].view(M, P)
        if lin:
            ls = float(host[5])
            r.aic = nacc * ls + 2 * (d + 1) * P
            r.bic = nacc * ls + math.log(nacc) * (d + 1) * P
        # host timeline of the call (us): entry -> enqueued, enqueued -> blocking read returned, read -> result built;
        # plus the gap since the previous call returned (the caller's own time)
        _t3 = time.perf_counter()
        hp = self.host_profile
        hp["calls"] += 1
        hp["enqueue_us"] += (_t1 - _t0) * 1e6
        hp["wait_us"] += (_t2 - _t1) * 1e6
        hp["result_us"] += (_t3 - _t2) * 1e6
        if self._t_last_return is not None:
            hp["between_calls_us"] += (_t0 - self._t_last_return) * 1e6
        self._t_last_return = _t3
        return r

    def _run(self, key, target, fn, graphable=True):
        """Run the device work ``fn(target) -> dict of result tensors`` of one call: eagerly the first time a key is
        seen, recorded as a CUDA graph the second time, replayed afterwards (one launch per call; the target is the only
        input that may change under a key - the tables are pinned by their addresses in the key).  -> (out, plan | None)

        Several ranks: whether to record is decided by the CALL COUNT of the key, which is the same on every rank as
        long as the ranks make the same calls; table addresses are rank-local, so they are left out of the history
        (a rank whose allocator returned a different address must not record one call later than its peers: with the
        NCCL fallback a recording rank would issue its collectives twice)."""
        graphable = graphable and self.use_graphs and self.timers is None and self.device.type == "cuda"
        if not graphable:
            return fn(target), None
        plan = self._graphs.get(key)
        if plan is not None:
            # the target is the only input that may change under a key.  It is copied on every replay: skipping the copy
            # for "the same tensor, same version counter" (-8 us) would silently serve a stale target whenever the caller
            # fills that tensor through a raw pointer (dlpack, cupy, an own kernel), which torch's counter does not see
            plan["target"].copy_(target)
            plan["gen"] += 1
            plan["graph"].replay()
            self.k.launches += plan["launches"]
            self._graphs[key] = self._graphs.pop(key)          # LRU order
            return plan["out"], plan
        hkey = key if self.comm.world == 1 else tuple(x for i, x in enumerate(key) if i not in self._PTR_SLOTS.get(key[0], ()))
        seen = self._seen.get(hkey, 0)
        if seen >= 1 and (self.comm.world == 1 or self.comm.peer is not None):
            plan = self._capture(key, target, fn)
            return plan["out"], plan
        if len(self._seen) > 256:
            self._seen.clear()
        self._seen[hkey] = seen + 1
        return fn(target), None

    _PTR_SLOTS = {"abc": (1, 4), "postpr": (1, 4), "columns": (1, 4)}   # positions of table addresses in the keys

    def _capture(self, key, target, fn):
        """Record the device work of this call as CUDA-graph segments (torch allocates its temporaries and outputs from
        a pool owned by the plan).  The tables are NOT referenced by the plan: the key pins their addresses."""
        while len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        tgt = target.clone()
        sg = _SegmentedGraph(self.device)
        l0 = self.k.launches
        torch.cuda.synchronize(self.device)
        sg.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(sg.stream):
            sg.begin()
            self.comm.recorder = sg
            try:
                out = fn(tgt)
            finally:
                self.comm.recorder = None
                sg.end()
        torch.cuda.current_stream(self.device).wait_stream(sg.stream)
        plan = dict(graph=sg, out=out, target=tgt, gen=0, launches=self.k.launches - l0,
                    keep=list(getattr(self.k, "_ws", {}).values()))   # workspaces must outlive the graphs
        self._graphs[key] = plan
        sg.replay()
        return plan

    def _abc_device(self, target, ss, par, pl, method, hcorr):
        """All device work of abc() — no host synchronisation — ending in the packed `small` buffer whose head is
        [status, n_acc, ds, region_const, all_mad_zero, sum log sigma2, 0, 0] (csrc/loclinear.cu:k_final_small)."""
        k = self.k
        d, n_local = ss.shape
        P = par.shape[0]
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        f = self._front(target, ss, status, pl, par)
        cap, idx, n_acc, ds, mad, dacc = f["cap"], f["idx"], f["n_acc"], f["ds"], f["mad"], f["dacc"]

        # The tail as three launches (two for rejection) with their reductions, NVLink exchanges and solves inside
        # (csrc/tail.cu); the stage-by-stage sequence below stays for d or P > 32 and for the NCCL fallback.
        use_tail = (self.fused_tail and hasattr(k, "tail_supported") and k.tail_supported(d, P)
                    and (self.comm.world == 1 or self.comm.peer is not None) and f["cand_par"] is None)
        if use_tail:
            lin = method == "loclinear"
            y, ss_out = k.new(P, cap), k.new(d, cap)
            xs = w = adj = res = None
            if lin:
                xs, w, adj, res = k.new(d, cap), k.new(cap), k.new(P, cap), k.new(P, cap)
            ws = k.tail_workspace(d, P)
            peer = self.comm.peer if self.comm.world > 1 else None
            with self._t("tail_gather_fit"):
                k.tail_gather_fit(peer, f["cand_ss"] if f["fused"] else None, f["slot"] if f["fused"] else None,
                                  f["perm"] if f["fused"] else None, None if f["fused"] else ss, par, dacc, idx, n_acc,
                                  f["row_offset"], cap, d, P, mad, target, ds, lin, xs, y, ss_out, w, ws, status)
            if lin and hcorr:
                with self._t("tail_resid_fit"):
                    k.tail_resid_fit(peer, xs, y, w, n_acc, cap, d, P, ws, status)
            small = k.new(k.final_small_len(d, P, lin))
            with self._t("tail_adjust_final"):
                k.tail_adjust_final(peer, xs, y, w, n_acc, ds, cap, d, P, lin, hcorr, mad, target, adj, res, ws, small,
                                    status)
            if self.comm.world > 1:
                nx = (2 + int(bool(hcorr))) if lin else 1
                self.comm.calls += nx
                self.comm.fused_calls += nx
            return dict(small=small, idx=idx, y=y, ss_out=ss_out, w=w, dacc=dacc, dist=f["dist"], adj=adj, res=res)

        xs, y, ss_out = k.new(d, cap), k.new(P, cap), k.new(d, cap)
        w = k.new(cap)
        row_major = par.stride(1) != 1          # (P, n) view of an n x P C-order array: see _as_table
        pcol, pld = (None, 0) if row_major else (par, par.stride(0))
        with self._t("gather"):
            if f["fused"]:   # the statistics of the accepted rows are already compact (candidate block)
                if f["cand_par"] is not None:
                    torch.cuda.current_stream(self.device).wait_stream(self._side)   # join the parameter pre-gather
                k.gather_cand(f["cand_ss"], f["cand_ss"].stride(0), f["slot"], f["perm"], pcol, pld, dacc, idx, n_acc,
                              f["row_offset"], cap, d, P, mad, target, ds, xs, y, ss_out, w, cand_par=f["cand_par"])
            else:
                k.gather(ss, ss.stride(0), pcol, pld, dacc, idx, n_acc, f["row_offset"], cap, d, P, mad,
                         target, ds, xs, y, ss_out, w)
            if row_major:    # one contiguous 8 P-byte record per accepted row instead of P scattered sectors
                k.gather_param_rows(par, par.stride(1), P, idx, n_acc, f["row_offset"], cap, y)
        # zero variance inside the accepted region (R cond2): local min / max now, reduced over the ranks at the end
        reg = k.new(d, 6)
        k.col_stats(ss_out, None, n_acc, cap, d, reg)

        beta = gamma = the_m = res = adj = sig = logsig = None
        stats = k.new(P, 6)
        _tl = self._t("loclinear")
        _tl.__enter__()
        if method == "loclinear":
            M = d + 1
            part = k.new(M * M + M * P)
            k.normal(xs, y, w, n_acc, cap, d, P, True, part)
            self.comm.allreduce_sum(part)
            gram, rhs = part[: M * M], part[M * M:]
            beta = k.new(M, P)
            k.solve(gram, rhs, d, P, beta, status)
            res = k.new(P, cap)
            # residuals, then sum res | sum w res | sum w res^2 in one statistics pass (measured faster than the fused
            # k_residual_stats, whose 48 running sums per thread cost occupancy: 0.15 vs 0.20 ms)
            k.residual(xs, y, beta, n_acc, cap, d, P, res)
            rs = k.new(P, 6)
            k.col_stats(res, w, n_acc, cap, P, rs)
            sums = k.new(3 * P + 1)                                 # sum res | sum w res | sum w res^2 | n_acc
            k.resid_moments(rs, n_acc, P, sums)
            self.comm.allreduce_sum(sums)
            # the_m = UNWEIGHTED mean of the residuals (R: colMeans), sigma2 around it, sum log sigma2 (aic / bic);
            # gram[0] = sum of the weights (already reduced over ranks)
            the_m, sig, logsig = k.new(P), k.new(P), k.new(1)
            k.resid_finish(sums, gram, P, the_m, sig, logsig)
            if hcorr:
                part2 = k.new(M * P)
                k.normal_logres(xs, res, the_m, w, n_acc, cap, d, P, part2)   # centres res in place
                self.comm.allreduce_sum(part2)
                gamma = k.new(M, P)
                k.solve(gram, part2, d, P, gamma, status)
            else:
                k.recentre_log(res, the_m, n_acc, cap, P, None)
            adj = k.new(P, cap)
            k.adjust(xs, res, beta, the_m, gamma, n_acc, cap, d, P, hcorr, adj)
            k.col_stats(adj, w, n_acc, cap, P, stats)
        else:
            k.col_stats(y, None, n_acc, cap, P, stats)
        _tl.__exit__()
        # Everything small goes into ONE buffer (host word | mad | stats | sigma2 | pred | coef | coef_h): one kernel, one
        # copy-out, one blocking read of its head.  Multi-rank: ONE min-exchange for everything that is a min, a max
        # (negated) or an OR (negated bits) - region min / max of the statistics, min / max of the posterior sample, the
        # status word - and one sum-exchange for the sums of the summary.
        pack = sm = None
        if self.comm.world > 1:
            pack, sm = k.new(2 * d + 2 * P + 8), k.new(P, 4)
            k.final_pack(status, reg, stats, d, P, pack, sm)
            self.comm.allreduce_min(pack)
            self.comm.allreduce_sum(sm)
        small = k.new(k.final_small_len(d, P, method == "loclinear"))
        k.final_small(status, n_acc, ds, reg, stats, mad, target, logsig, beta, gamma, the_m, sig, pack, sm, d, P, small)
        return dict(small=small, idx=idx, y=y, ss_out=ss_out, w=w, dacc=dacc, dist=f["dist"], adj=adj, res=res)

    def abc_columns(self, target, param, sumstat, tol, method, hcorr: bool = True, shard=None) -> List[AbcResult]:
        """ABC-DLS's per-column mode (ABC.py:1949-1953, 2805-2809): P separate 1-D abc() problems,
        column c of ``sumstat`` against column c of ``param``."""
        ss, par = _as_table(sumstat), _as_table(param, row_major_ok=True)
        target = target.reshape(-1)
        return [self.abc(target[c:c + 1], par[c:c + 1], ss[c:c + 1], tol, method, hcorr, shard=shard)
                for c in range(par.shape[0])]

    def abc_columns_summary(self, target, param, sumstat, tol, method, hcorr: bool = True, shard=None,
                            want=("min", "max", "median", "mean")):
        """Per-column mode when only ``summary(res)`` is kept - ABC-DLS's SMC reads Min and Max of every column
        (ABC.py:2805-2840): {"min", "max", "median", "mean"} -> (P,) float64 numpy arrays (those named in ``want``).
        One rank and a table that fits a CTA's shared memory: the P problems are ONE launch of csrc/small.cu.  Large
        tables, any number of ranks: ONE pass over the N x P statistics for all P problems (csrc/columns.cu).
        Otherwise a loop over abc()."""
        import numpy as np
        from . import cv as _cv, summary as _summary
        ss, par = _as_table(sumstat), _as_table(param, row_major_ok=True)
        target = target.to(device=self.device, dtype=torch.float64).reshape(-1)
        P = par.shape[0]
        if ss.shape[0] != P or target.numel() != P:
            raise AbcError("per-column mode needs one summary statistic and one target value per parameter")
        if method not in ("rejection", "loclinear"):
            raise AbcError("Method must be rejection or loclinear (neuralnet / ridge stay with R's nnet)")
        o = None
        if self.comm.world == 1 and ss.device == self.device and par.device == self.device:
            o = _cv.small_fits(self, par, ss, np.arange(P, dtype=np.int32), None, float(tol), method, hcorr, target=target)
        if o is not None:
            full = {"median": o[:, 0].copy(), "mean": o[:, 1].copy(), "min": o[:, 4].copy(), "max": o[:, 5].copy(),
                    "ds": o[:, 2].copy(), "mad": o[:, 3].copy()}
            return {k_: full[k_] for k_ in tuple(want) + ("ds", "mad")}
        got = self._columns_streaming(target, par, ss, tol, method, hcorr, shard, want)
        if got is not None:
            return got
        out = {k_: np.empty(P) for k_ in tuple(want) + ("ds", "mad")}
        # (no shard= here: a one-column slice of the table need not have the aligned layout that promise is about)
        for c, r in enumerate(self.abc_columns(target, par, ss, tol, method, hcorr)):
            w = r.adj_values is not None
            out["ds"][c], out["mad"][c] = r.ds, float(r.mad[0])
            if "min" in out:
                out["min"][c] = float(r.stats[0, 0])
            if "max" in out:
                out["max"][c] = float(r.stats[0, 1])
            if "mean" in out:       # stats are already global: sum w x / sum w
                out["mean"][c] = float(r.stats[0, 3] / r.stats[0, 5]) if w else float(r.stats[0, 2] / r.nacc)
            if "median" in out:
                vals, wts = r.values, r.weights
                if self.comm.world > 1:
                    vals, wts = self.gather_sample(vals, wts)
                out["median"][c] = float(_summary.point_estimate(vals, wts, r.stats, w, "median")[0])
        return out

    # ---- per-column mode as ONE pass over the N x P statistics (csrc/columns.cuh) ---------------------------------------
    def _columns_streaming(self, target, par, ss, tol, method, hcorr, shard, want):
        """All P one-dimensional rejection problems from one read of the statistics table; any number of ranks.
        None when the table does not qualify or the device flags a bracket miss (the caller then loops over abc())."""
        import numpy as np
        from . import summary as _summary
        k = self.k
        if method != "rejection" or not hasattr(k, "columns_supported") or not self.use_columns:
            return None
        P, n_local = ss.shape
        ok = (ss.device == self.device and ss.stride(1) == 1 and k.columns_supported(ss, n_local, P, ss.stride(0))
              and (par.device == self.device or (par.device.type == "cpu" and par.is_pinned()))
              and (self.comm.world == 1 or self.comm.peer is not None))
        if shard is not None and self.comm.world > 1:
            n_global, row_offset = int(shard[0]), int(shard[1])
            per = -(-n_global // self.comm.world)
            ok = ok and min(per, n_global - (self.comm.world - 1) * per) >= self.FAST_MIN_ROWS
        else:
            n_global, row_offset, ns, ok = self._shard_info(n_local, int(ok))
        if not ok:
            return None
        nacc = int(math.ceil(n_global * tol))
        if not (1 <= nacc <= n_global):
            raise AbcError("'tol' must give 1 <= ceiling(N * tol) <= N")
        hint = -(-nacc * n_local // n_global)
        lists = "median" in want
        key = ("columns", ss.data_ptr(), tuple(ss.shape), ss.stride(0), par.data_ptr(), par.stride(0), par.stride(1),
               par.device.type, float(tol), self.comm.world, n_global, row_offset, lists)
        _t0 = time.perf_counter()
        o, plan = self._run(key, target, lambda tgt: self._columns_device(tgt, par, ss, n_global, nacc, hint, lists))
        self.host_enqueue_ms.append((time.perf_counter() - _t0) * 1e3)
        small = o["small"].cpu()                    # the one blocking read
        st = int(small[0])
        if st & _capi.S_PEER_TIMEOUT:
            raise RuntimeError("a rank did not take part in an exchange of this call (NVLink mailbox wait timed out)")
        per_col = small[8:].view(P, 8).numpy()
        if (st & _capi.S_FASTPATH_MISS) or (per_col[:, 1] == 0).any() or (per_col[:, 5] != nacc).any():
            self.fast_misses += 1                   # a zero MAD needs R's zero-variance checks: the exact path has them
            return None
        if st & _capi.S_NONFINITE:
            raise AbcError("non-finite values in 'sumstat'/'target' (R's NA-row handling is not implemented)")
        self.path_calls["columns"] = self.path_calls.get("columns", 0) + 1
        for c in np.flatnonzero(per_col[:, 7] != 0):
            print("Warning message:\nZero variance in the summary statistics in the selected region. Check summary "
                  f"statistics, consider larger tolerance. (column {c + 1})")
        out = {}
        if "min" in want:
            out["min"] = per_col[:, 2].copy()
        if "max" in want:
            out["max"] = per_col[:, 3].copy()
        if "mean" in want:
            out["mean"] = per_col[:, 4] / per_col[:, 5]
        if lists:
            cnts = o["acc_cnt"].cpu().tolist()
            med = np.empty(P)
            for c in range(P):
                vals = o["acc_val"][c, : cnts[c]].reshape(-1, 1)
                vals, _ = self.gather_sample(vals, torch.ones(vals.shape[0], dtype=torch.float64, device=self.device))
                med[c] = float(_summary.point_estimate(vals, None, None, False, "median")[0])
            out["median"] = med
        out["ds"], out["mad"] = per_col[:, 0].copy(), per_col[:, 1].copy()
        return {k_: out[k_] for k_ in tuple(want) + ("ds", "mad")}

    def _columns_device(self, target, par, ss, n_global, nacc, hint, lists):
        k = self.k
        P, n_local = ss.shape
        comm = self.comm if self.comm.world > 1 else None
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        med, mad = k.new(P), k.new(P)
        with self._t("columns_front"):
            ws, dist, cnt = k.columns_front(ss, n_local, n_global, P, ss.stride(0), target, nacc, hint, med, mad, status,
                                            comm)
        with self._t("columns_select"):
            sel = self._select(dist, dist.shape[1], None, nacc - 1, nacc - 1, counts=cnt)      # (P, 2): ds_j twice
        with self._t("columns_accept"):
            c2 = k.columns_count(ws, n_local, P, hint, nacc, sel, mad, target, cnt, status)
            # R's cap rule: the first nacc rows, in row order, of {dist <= ds}; lower ranks hold the lower rows
            mine = c2.sum(dim=1)
            before = torch.zeros_like(mine)
            if comm is not None:
                before = comm.allgather(mine.contiguous())[: comm.rank].sum(0)
            quota = torch.minimum((nacc - before).clamp(min=0), mine)
            keys = k.columns_rowkeys(ws, n_local, P, hint, sel, cnt, dist.shape[1])
            k0 = (quota - 1).clamp(min=0).contiguous()
            thr = self._select(keys, keys.shape[1], None, k0, k0, counts=cnt, local=True)          # (P, 2)
            thr = torch.where((quota > 0)[:, None], thr, torch.full_like(thr, -1.0)).contiguous()
            small = k.new(8 + 8 * P)
            acc = None
            if lists:
                cap = min(n_local, 2 * hint + 1024)
                acc = (cap, torch.empty((P, cap), dtype=torch.int32, device=self.device), k.new(P, cap), k.new(P, cap),
                       torch.zeros(P, dtype=torch.int64, device=self.device))
            k.columns_accept(self.comm.peer if comm is not None else None, ws, n_local, P, hint, sel, med, mad, cnt,
                             thr, par, small, status, acc)
            if comm is not None:
                self.comm.calls += 1
                self.comm.fused_calls += 1
        out = dict(small=small)
        if lists:
            out.update(acc_row=acc[1], acc_val=acc[2], acc_dist=acc[3], acc_cnt=acc[4])
        return out

    def gather_sample(self, vals: torch.Tensor, w: torch.Tensor):
        """The posterior sample of ALL ranks (rank order = row order) on every rank: (n_acc, P) values, (n_acc,) weights."""
        c = self.comm
        if c.world == 1:
            return vals, w
        n_loc = torch.tensor([vals.shape[0]], dtype=torch.int64, device=vals.device)
        ns = c.allgather(n_loc).view(-1).cpu().tolist()
        cap = max(max(ns), 1)
        pv = torch.zeros((cap, vals.shape[1]), dtype=torch.float64, device=vals.device)
        pw = torch.zeros(cap, dtype=torch.float64, device=vals.device)
        pv[: vals.shape[0]], pw[: vals.shape[0]] = vals, w
        gv, gw = c.allgather(pv), c.allgather(pw)
        return (torch.cat([gv[i, : ns[i]] for i in range(c.world)]), torch.cat([gw[i, : ns[i]] for i in range(c.world)]))

    # ---- postpr() -------------------------------------------------------------------------------
    def postpr(self, target: torch.Tensor, model_id: torch.Tensor, models: List[str], sumstat: torch.Tensor,
               tol: float, method: str = "rejection", keep_dist: bool = False, shard=None,
               _level: int = 0) -> PostprResult:
        """``abc::postpr(method='rejection')``: model_id int32 (n_local,) indexes ``models`` (sorted levels, the same
        list on every rank).  Same footing as abc(): one packed small buffer, the status OR-ed and the counts summed
        over the ranks on the device, ONE blocking read, replayed as a CUDA graph from the third call on."""
        if method != "rejection":
            raise AbcError("postpr: only method='rejection' is implemented (mnlogistic / neuralnet stay with R's nnet)")
        k = self.k
        _t0 = time.perf_counter()
        if not (target.dtype == torch.float64 and target.device == self.device and target.dim() == 1
                and target.is_contiguous()):
            target = target.to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
        K = len(models)
        # same memo as abc(): a repeated call on the same table objects skips validation, plan and key
        sig = (id(sumstat), id(model_id), sumstat.data_ptr(), model_id.data_ptr(), tuple(sumstat.shape), sumstat.stride(),
               K, float(tol), bool(keep_dist), int(_level), None if shard is None else (int(shard[0]), int(shard[1])))
        memo = self._last_postpr if (self.comm.world == 1 or shard is not None) else None   # see abc()
        if memo is not None and memo[0] == sig and memo[1]() is sumstat and memo[2]() is model_id:
            ss = sumstat
            _, _, _, d, n_local, pl, key = memo
            if target.numel() != d:
                raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
        else:
            ss = _as_table(sumstat)
            d, n_local = ss.shape
            if K < 2:
                raise AbcError("At least two different models must be given.")
            if target.numel() != d:
                raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
            if model_id.numel() != n_local:
                raise AbcError("'index' must be the same length as the number of rows in 'sumstat'.")
            mid0 = model_id
            model_id = model_id.to(device=self.device, dtype=torch.int32).contiguous()
            pl = self._plan(ss, tol, keep_dist, _level, shard)
            key = ("postpr", ss.data_ptr(), tuple(ss.shape), ss.stride(0), model_id.data_ptr(), K, float(tol),
                   bool(keep_dist), int(_level), self.comm.world, pl["n_global"], pl["row_offset"])
            self._last_postpr = ((sig, weakref.ref(sumstat), weakref.ref(mid0), d, n_local, pl, key)
                                 if (ss is sumstat and model_id is mid0 and ss.device == self.device) else None)
        self.path_calls[pl["path"]] += 1
        o, plan = self._run(key, target, lambda tgt: self._postpr_device(tgt, ss, model_id, K, pl),
                            graphable=not keep_dist)
        self.host_enqueue_ms.append((time.perf_counter() - _t0) * 1e3)
        small = o["small"].cpu().tolist()             # the one blocking read
        st, n_loc, ds_h = int(small[0]), int(small[1]), float(small[2])
        if (st & _capi.S_FASTPATH_MISS) and pl["fast"]:
            self.fast_misses += 1
            return self.postpr(target, model_id, models, sumstat, tol, method, keep_dist, shard,
                               _level=1 if pl["fused"] else 2)
        if st & _capi.S_PEER_TIMEOUT:
            raise RuntimeError("a rank did not take part in an exchange of this call (NVLink mailbox wait timed out)")
        if st & _capi.S_NONFINITE:
            raise AbcError("non-finite values in 'sumstat'/'target' (R's NA-row handling is not implemented)")
        if small[3] != 0 and self._zero_variance(ss):
            raise AbcError("Zero variance in the summary statistics.")
        if st & _capi.S_CAPACITY:
            raise AbcError("internal: accepted-row buffer overflow")
        # counts / proportions / MADs come from the host copy of the small buffer, built on first access (the host time
        # spent here delays the next call); the accepted rows are copied out of the graph's buffers now
        dev, cnt_h, mad_h = self.device, [int(c) for c in small[8 + d: 8 + d + K]], small[8:8 + d]
        tot = float(sum(cnt_h))
        idx = o["idx"][:n_loc].clone() if plan is not None else o["idx"][:n_loc]
        row_offset = pl["row_offset"]
        return PostprResult(models=list(models), counts=_Lazy(lambda: torch.tensor(cnt_h, dtype=torch.int64, device=dev)),
                            pred=_Lazy(lambda: torch.tensor([c / tot for c in cnt_h], dtype=torch.float64, device=dev)),
                            nacc=pl["nacc"], n_global=pl["n_global"], ds=ds_h,
                            mad=_Lazy(lambda: torch.tensor(mad_h, dtype=torch.float64, device=dev)),
                            idx=idx, ss=lambda: ss[:, idx - row_offset].t(), values=lambda: model_id[idx - row_offset],
                            dist=o["dist"] if keep_dist else None)

    def _postpr_device(self, target, ss, model_id, K, pl):
        """All device work of postpr() ending in small = [status, n_acc, ds, every MAD zero, 0 x 4 | mad | counts]."""
        k = self.k
        d = ss.shape[0]
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        f = self._front(target, ss, status, pl)
        counts = torch.zeros(K, dtype=torch.int64, device=self.device)
        k.postpr_counts(model_id, f["idx"], f["n_acc"], f["row_offset"], f["cap"], K, counts)
        small = k.new(8 + d + K)
        if hasattr(k, "postpr_final") and (self.comm.world == 1 or self.comm.peer is not None):
            k.postpr_final(self.comm.peer if self.comm.world > 1 else None, status, f["n_acc"], f["ds"], f["mad"], d,
                           counts, K, small)
            if self.comm.world > 1:
                self.comm.calls += 1
                self.comm.fused_calls += 1
        else:   # NCCL fallback / CPU stage double: the same buffer from framework ops
            self.comm.allreduce_sum(counts)
            bits = torch.stack([(status[0] >> b) & 1 for b in range(8)]).to(torch.int64)
            self.comm.allreduce_max(bits)
            st = (bits << torch.arange(8, device=bits.device)).sum().to(torch.float64).view(1)
            small[:8] = torch.cat([st, f["n_acc"].to(torch.float64), f["ds"].view(1),
                                   (f["mad"] == 0).all().to(torch.float64).view(1), small.new_zeros(4)])
            small[8:8 + d] = f["mad"]
            small[8 + d:] = counts.to(torch.float64)
        return dict(small=small, idx=f["idx"], dist=f["dist"])

    # ---- SMC table passes -------------------------------------------------------------------------
    def oldrange(self, param_all: torch.Tensor):
        """Per-column (min, max) over all simulated parameter rows, all ranks (ABC.py:2714-2716)."""
        tab = _as_table(param_all)
        P, n = tab.shape
        mm = self.k.new(2 * P)
        self.k.smc_oldrange(tab, n, P, tab.stride(0), mm)
        mn, mx = mm[:P].clone(), mm[P:].clone()
        self.comm.allreduce_min(mn)
        self.comm.allreduce_max(mx)
        return mn, mx

    def box_filter(self, param_all: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, shard=None) -> torch.Tensor:
        """Ascending GLOBAL row numbers of this rank's rows with every parameter in [lo, hi] (ABC.py:2974-2990).
        shard = (n_global, row_offset) saves the blocking exchange of the shard sizes."""
        tab = _as_table(param_all)
        P, n = tab.shape
        row_offset = int(shard[1]) if shard is not None else self._shard_info(n)[1]
        lo = lo.to(device=self.device, dtype=torch.float64).contiguous()
        hi = hi.to(device=self.device, dtype=torch.float64).contiguous()
        idx = self.k.new(n, dtype=torch.int64)
        n_out = torch.zeros(1, dtype=torch.int64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.k.smc_box_filter(tab, n, P, tab.stride(0), lo, hi, row_offset, n, idx, n_out, status)
        return idx[: int(n_out.item())]
