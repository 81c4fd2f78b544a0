"""CUDA stage layer: one method per C-ABI entry point, taking torch CUDA tensors.

This is the only place that marshals tensors into ``libabcdls.so`` (pointers + sizes + the current
CUDA stream).  ``engine.py`` sequences these stages and puts the NCCL collectives between them.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class CudaStages:
    """Thin, stateless-but-for-workspaces wrapper over the C ABI on one device."""

    def __init__(self, device: torch.device):
        if not torch.cuda.is_available():
            raise RuntimeError("abc_dls_b200 needs a CUDA (sm_100a / B200) device; there is no CPU path")
        self.device = torch.device(device)
        self.handle = _capi.Handle(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.lib = self.handle.lib
        self._ws = {}
        self.launches = 0  # kernels enqueued through this object (bench.py reports it)
        self.timer = None  # optional: callable(name) -> context manager recording CUDA events (engine._t)
        self._side = None  # second stream for work that only needs the pass (candidate-row compaction)

    def _timed(self, name):
        import contextlib
        return self.timer(name) if self.timer is not None else contextlib.nullcontext()

    # ---- helpers ----------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, key: str, nbytes: int) -> torch.Tensor:
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    def new(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ---- selection ----------------------------------------------------------------------------
    def select_workspace(self, njobs: int) -> torch.Tensor:
        return self.workspace(f"select{njobs}", self.lib.abcdls_select_workspace_bytes(njobs))

    def select_begin(self, ws, njobs, base, n, ld, counts, center, k0, k1):
        k0d = k0 if torch.is_tensor(k0) else None
        k1d = k1 if torch.is_tensor(k1) else None
        self.handle.call("abcdls_select_begin", _p(ws), ws.numel(), njobs, _p(base), n, ld, _p(counts), _p(center),
                         _p(k0d), _p(k1d), 0 if k0d is not None else int(k0), 0 if k1d is not None else int(k1),
                         self._stream())
        self.launches += 2

    def select_hist(self, ws, njobs, n_max, p):
        self.handle.call("abcdls_select_hist", _p(ws), njobs, n_max, p, self._stream())
        self.launches += 1

    def select_pick(self, ws, njobs, p):
        self.handle.call("abcdls_select_pick", _p(ws), njobs, p, self._stream())
        self.launches += 1

    def select_hist_tensor(self, ws, njobs) -> torch.Tensor:
        buf, cnt = C.c_void_p(), C.c_size_t()
        self.handle.call("abcdls_select_hist_buffer", _p(ws), njobs, C.byref(buf), C.byref(cnt))
        off = buf.value - ws.data_ptr()
        return ws[off: off + cnt.value * 8].view(torch.int64)

    def select_end(self, ws, njobs, out):
        self.handle.call("abcdls_select_end", _p(ws), njobs, _p(out), self._stream())
        self.launches += 1

    def median_from_ranks(self, ranks2, ncols, scale, out):
        self.handle.call("abcdls_median_from_ranks", _p(ranks2), ncols, float(scale), _p(out), self._stream())
        self.launches += 1

    # ---- kernel 1 -------------------------------------------------------------------------------
    def scale_dist(self, ss, n, d, ld, mad, target, dist, status):
        self.handle.call("abcdls_scale_dist", _p(ss), n, d, ld, _p(mad), _p(target), _p(dist), _p(status),
                         self._stream())
        self.launches += 1

    # ---- kernel 2 -------------------------------------------------------------------------------
    def accept_workspace(self, n):
        return self.workspace("accept", self.lib.abcdls_accept_workspace_bytes(n))

    def count_le(self, dist, n, n_dev, ds, le_count, ws):
        self.handle.call("abcdls_count_le", _p(dist), n, _p(n_dev), _p(ds), _p(le_count), _p(ws), ws.numel(),
                         self._stream())
        self.launches += 2        # k_tile_count_le, k_scan_tiles

    def accept(self, dist, n, n_dev, src_idx, ds, quota, row_offset, cap, idx, dist_out, n_out, ws, status,
               slot_out=None):
        self.handle.call("abcdls_accept", _p(dist), n, _p(n_dev), _p(src_idx), _p(ds), _p(quota), row_offset, cap,
                         _p(idx), _p(dist_out), _p(slot_out), _p(n_out), _p(ws), ws.numel(), _p(status), self._stream())
        self.launches += 1

    # ---- one-pass front half (csrc/fused.cuh) -----------------------------------------------------
    def fused_supported(self, ss, n, d, ld) -> bool:
        return bool(self.lib.abcdls_fused_supported(_p(ss), n, d, ld))

    def fused_emit_cap(self, cap) -> int:
        return int(self.lib.abcdls_fused_emit_cap(cap))

    def fused_emit_pitch(self, d) -> int:
        return int(self.lib.abcdls_fused_emit_pitch(int(d)))

    def fused_front(self, ss, n, n_global, d, ld, target, nacc, row_offset, cap, med, mad, cand_row, cand_dist, cand_ss,
                    cand_lrow, perm, n_cand, status, comm=None, pregather=None, fine=0):
        """Medians + MADs AND the candidate rows of the tolerance cut from ONE read of the table.  Leaves the exact
        distance of every candidate in cand_dist (row order); the caller selects ds among them (all ranks) and then
        calls fused_check."""
        world = comm.world if comm is not None else 1
        ws = self.workspace("fused", self.lib.abcdls_fused_workspace_bytes(n, d))
        p0, c0, p1, c1 = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        st = self._stream()
        self.handle.call("abcdls_fused_sample_cols", _p(ss), n, d, ld, world, fine, _p(ws), ws.numel(), C.byref(p0),
                         C.byref(c0), st)
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.float64, 8))
        if fine:
            # second-stage counting sample against the first-stage windows (narrower brackets for the pass)
            with self._timed("k_sample_fine"):
                self.handle.call("abcdls_fused_sample_fine", _p(ss), n, d, ld, world, fine, _p(ws), ws.numel(),
                                 C.byref(p0), C.byref(c0), st)
            if world > 1:
                comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
            self.launches += 3
        self.handle.call("abcdls_fused_sample_dist", _p(ss), n, d, ld, _p(target), nacc, n_global, world, fine, _p(ws),
                         ws.numel(), C.byref(p0), C.byref(c0), st)
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.float64, 8))
        with self._timed("k_fused_pass"):
            self.handle.call("abcdls_fused_pass", _p(ss), n, d, ld, world, fine, _p(cand_ss), _p(cand_lrow), cand_ss.shape[0],
                             _p(ws), ws.numel(), _p(status), C.byref(p0), C.byref(c0), st)
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
        # the candidate rows need only the pass: compact them on a second stream (three small kernels) while this one
        # refines medians / MADs; joined before the exact distances.  (NCCL fallback: segments cannot span a fork.)
        fork = world == 1 or getattr(comm, "peer", None) is not None
        main = torch.cuda.current_stream(self.device)
        if fork:
            if self._side is None:
                self._side = torch.cuda.Stream(self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                self.handle.call("abcdls_fused_rows", _p(ss), n, d, ld, world, row_offset, cap, _p(cand_row), _p(n_cand),
                                 _p(ws), ws.numel(), _p(status), C.c_void_p(self._side.cuda_stream))
        else:
            self.handle.call("abcdls_fused_rows", _p(ss), n, d, ld, world, row_offset, cap, _p(cand_row), _p(n_cand),
                             _p(ws), ws.numel(), _p(status), st)
        gs = gc = None
        with self._timed("fused_refine"):
            for phase in range(4):
                self.handle.call("abcdls_fused_refine", _p(ss), n, n_global, d, ld, phase, world, _p(gs), _p(gc), _p(ws),
                                 ws.numel(), _p(med), _p(mad), _p(status), C.byref(p0), C.byref(c0), C.byref(p1),
                                 C.byref(c1), st)
                if world > 1:
                    if phase in (0, 2):
                        gs, gc = comm.allgather_ragged(self._view(ws, p0, c0, torch.float64, 8),
                                                       self._view(ws, p1, c1, torch.int32, 4))
                    elif phase == 1:
                        comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
        if fork:
            main.wait_stream(self._side)
        with self._timed("fused_candidates"):
            self.handle.call("abcdls_fused_exact", _p(ss), n, d, ld, world, _p(mad), _p(target), cap, _p(cand_dist),
                             _p(cand_ss), _p(cand_lrow), cand_ss.shape[0], _p(perm), _p(ws), ws.numel(), _p(status), st)
        self.launches += 22       # the kernels of profiles/r02_launches_1e8_final.txt up to k_cand_exact (+ 3 above with `fine`)

    def fused_ds(self, n, d, nacc, cap, cand_dist, n_cand, ds, status, comm=None):
        """Exact ds = nacc-th smallest candidate distance over all ranks (3 small kernels-stages, 2 exchanges)."""
        world = comm.world if comm is not None else 1
        ws = self._ws["fused"]
        p0, c0 = C.c_void_p(), C.c_size_t()
        gath = None
        for phase in range(3):
            self.handle.call("abcdls_fused_ds", n, d, phase, world, _p(gath), _p(cand_dist), _p(n_cand), cap, nacc,
                             _p(ds), _p(ws), ws.numel(), _p(status), C.byref(p0), C.byref(c0), self._stream())
            if world > 1 and phase == 0:
                comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
            elif world > 1 and phase == 1:
                gath = comm.allgather(self._view(ws, p0, c0, torch.float64, 8)).contiguous()
        self.launches += 6        # plan, hist, pick, gather, pack, final

    def fused_check(self, n, d, mad, ds, status):
        ws = self._ws["fused"]
        self.handle.call("abcdls_fused_check", n, d, _p(mad), _p(ds), _p(ws), ws.numel(), _p(status), self._stream())
        self.launches += 1

    def gather_cand(self, cand_ss, ld_cand, slot, perm, param, ld_param, dist_acc, idx, n_acc, row_offset, cap, d, P,
                    mad, target, ds, xs, y, ss_out, w, cand_par=None):
        self.handle.call("abcdls_gather_cand", _p(cand_ss), ld_cand, _p(slot), _p(perm), _p(param), ld_param,
                         _p(cand_par), cand_par.stride(0) if cand_par is not None else 0, _p(dist_acc),
                         _p(idx), _p(n_acc), row_offset, cap, d, P, _p(mad), _p(target), _p(ds), _p(xs), _p(y),
                         _p(ss_out), _p(w), self._stream())
        self.launches += 1

    # ---- fast path (sampled brackets) -----------------------------------------------------------
    def _buf(self, ws, ptr, count, dtype):
        off = ptr.value - ws.data_ptr()
        return ws[off: off + count.value * 8].view(dtype)

    def _view(self, ws, ptr, count, dtype, itemsize):
        off = ptr.value - ws.data_ptr()
        return ws[off: off + count.value * itemsize].view(dtype)

    def mad_fast(self, ss, n, n_global, d, ld, med, mad, status, comm=None):
        """Exact median + MAD of every column with ONE table read.  comm (world > 1): all-reduce / all-gather
        between the stages (csrc/bracket.cu phase comments)."""
        world = comm.world if comm is not None else 1
        ws = self.workspace("madfast", self.lib.abcdls_mad_fast_workspace_bytes(n, d))
        p0, c0, p1, c1 = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        self.handle.call("abcdls_mad_fast_sample", _p(ss), n, d, ld, world, _p(ws), ws.numel(), C.byref(p0),
                         C.byref(c0), self._stream())
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.float64, 8))
        with self._timed("k_bracket_pass"):
            self.handle.call("abcdls_mad_fast_pass", _p(ss), n, d, ld, world, _p(ws), ws.numel(), C.byref(p0),
                             C.byref(c0), self._stream())
        if world == 1:
            self.handle.call("abcdls_mad_fast_finish", n, n_global, d, _p(ws), ws.numel(), _p(med), _p(mad),
                             _p(status), self._stream())
            self.launches += 15
            return
        comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
        gs = gc = None
        for phase in range(4):
            self.handle.call("abcdls_mad_fast_refine", n, n_global, d, phase, world, _p(gs), _p(gc), _p(ws), ws.numel(),
                             _p(med), _p(mad), _p(status), C.byref(p0), C.byref(c0), C.byref(p1), C.byref(c1),
                             self._stream())
            if phase in (0, 2):
                gs, gc = comm.allgather_ragged(self._view(ws, p0, c0, torch.float64, 8),
                                               self._view(ws, p1, c1, torch.int32, 4))
            elif phase == 1:
                comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
        self.launches += 15

    def dist_fast(self, ss, n, n_global, d, ld, mad, target, nacc, row_offset, cap, dist, cand_row, cand_dist,
                  n_cand, ds, status, comm=None):
        """Fused distance + tolerance bracket: candidate rows in row order and the exact global threshold ds."""
        world = comm.world if comm is not None else 1
        ws = self.workspace("distfast", self.lib.abcdls_dist_fast_workspace_bytes(n, cap))
        p0, c0, p1, c1 = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        self.handle.call("abcdls_dist_fast_sample", _p(ss), n, d, ld, _p(mad), _p(target), nacc, n_global, world, cap,
                         _p(ws), ws.numel(), C.byref(p0), C.byref(c0), self._stream())
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.float64, 8))
        with self._timed("k_dist_pass"):
            self.handle.call("abcdls_dist_fast_pass", _p(ss), n, d, ld, _p(mad), _p(target), nacc, world, row_offset,
                             cap, _p(dist), _p(cand_row), _p(cand_dist), _p(n_cand), _p(ds), _p(ws), ws.numel(),
                             _p(status), C.byref(p0), C.byref(c0), self._stream())
        self.launches += 12
        if world == 1:
            return
        comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
        self.handle.call("abcdls_dist_fast_refine", n, nacc, cap, 0, world, None, None, _p(cand_dist), _p(ds), _p(ws),
                         ws.numel(), _p(status), C.byref(p0), C.byref(c0), C.byref(p1), C.byref(c1), self._stream())
        gs = comm.allgather(self._view(ws, p0, c0, torch.float64, 8)).contiguous()
        gc = comm.allgather(self._view(ws, p1, c1, torch.int32, 4)).contiguous()
        self.handle.call("abcdls_dist_fast_refine", n, nacc, cap, 1, world, _p(gs), _p(gc), _p(cand_dist), _p(ds),
                         _p(ws), ws.numel(), _p(status), C.byref(p0), C.byref(c0), C.byref(p1), C.byref(c1),
                         self._stream())

    # ---- kernel 3 -------------------------------------------------------------------------------
    def gather(self, ss, ld_ss, param, ld_param, dist_acc, idx, n_acc, row_offset, cap, d, P, mad, target, ds, xs, y,
               ss_out, w):
        self.handle.call("abcdls_gather", _p(ss), ld_ss, _p(param), ld_param, _p(dist_acc), _p(idx), _p(n_acc),
                         row_offset, cap, d, P, _p(mad), _p(target), _p(ds), _p(xs), _p(y), _p(ss_out), _p(w),
                         self._stream())
        self.launches += 1

    def gather_param_rows(self, param, row_pitch, P, idx, n_acc, row_offset, cap, y):
        """Parameter half of a gather for a ROW-major table (device or pinned host)."""
        self.handle.call("abcdls_gather_param_rows", _p(param), row_pitch, P, _p(idx), _p(n_acc), row_offset, cap, _p(y),
                         self._stream())
        self.launches += 1

    def normal(self, xs, rhs, w, n_acc, cap, d, P, with_gram, partial):
        ws = self.workspace("normal", self.lib.abcdls_normal_workspace_bytes(d, P))
        self.handle.call("abcdls_normal", _p(xs), _p(rhs), _p(w), _p(n_acc), cap, d, P, int(with_gram), _p(partial),
                         _p(ws), ws.numel(), self._stream())
        self.launches += 2

    def solve(self, gram, rhs, d, P, beta, status):
        self.handle.call("abcdls_solve", _p(gram), _p(rhs), d, P, _p(beta), _p(status), self._stream())
        self.launches += 1

    def residual(self, xs, y, beta, n_acc, cap, d, P, res):
        self.handle.call("abcdls_residual", _p(xs), _p(y), _p(beta), _p(n_acc), cap, d, P, _p(res), self._stream())
        self.launches += 1

    def residual_stats(self, xs, y, beta, w, n_acc, cap, d, P, res, stat):
        key = f"resid{P}"
        if key not in self._ws:
            self.workspace(key, self.lib.abcdls_residual_workspace_bytes(P)).zero_()
        ws = self._ws[key]
        self.handle.call("abcdls_residual_stats", _p(xs), _p(y), _p(beta), _p(w), _p(n_acc), cap, d, P, _p(res),
                         _p(stat), _p(ws), ws.numel(), self._stream())
        self.launches += 1

    def normal_logres(self, xs, res, the_m, w, n_acc, cap, d, P, partial):
        ws = self.workspace("normal", self.lib.abcdls_normal_workspace_bytes(d, P))
        self.handle.call("abcdls_normal_logres", _p(xs), _p(res), _p(the_m), _p(w), _p(n_acc), cap, d, P, _p(partial),
                         _p(ws), ws.numel(), self._stream())
        self.launches += 2

    def recentre_log(self, res, the_m, n_acc, cap, P, logres2):
        self.handle.call("abcdls_recentre_log", _p(res), _p(the_m), _p(n_acc), cap, P, _p(logres2), self._stream())
        self.launches += 1

    def adjust(self, xs, res, beta, the_m, gamma, n_acc, cap, d, P, hcorr, adj):
        self.handle.call("abcdls_adjust", _p(xs), _p(res), _p(beta), _p(the_m), _p(gamma), _p(n_acc), cap, d, P,
                         int(hcorr), _p(adj), self._stream())
        self.launches += 1

    def col_stats(self, values, w, n_acc, cap, ncols, out):
        key = f"colstats{ncols}"
        if key not in self._ws:
            self.workspace(key, self.lib.abcdls_col_stats_workspace_bytes(ncols)).zero_()
        ws = self._ws[key]
        self.handle.call("abcdls_col_stats", _p(values), _p(w), _p(n_acc), cap, ncols, _p(out), _p(ws), ws.numel(),
                         self._stream())
        self.launches += 1

    def resid_moments(self, rs, n_acc, P, sums):
        self.handle.call("abcdls_resid_moments", _p(rs), _p(n_acc), P, _p(sums), self._stream())
        self.launches += 1

    def resid_finish(self, sums, gram, P, the_m, sig, logsig):
        self.handle.call("abcdls_resid_finish", _p(sums), _p(gram), P, _p(the_m), _p(sig), _p(logsig), self._stream())
        self.launches += 1

    def final_small_len(self, d, P, loclinear) -> int:
        return int(self.lib.abcdls_final_small_doubles(d, P, int(loclinear)))

    def final_pack(self, status, reg, stats, d, P, pack, sm):
        self.handle.call("abcdls_final_pack", _p(status), _p(reg), _p(stats), d, P, _p(pack), _p(sm), self._stream())
        self.launches += 1

    def final_small(self, status, n_acc, ds, reg, stats, mad, target, logsig, beta, gamma, the_m, sig, pack, sm, d, P,
                    small):
        self.handle.call("abcdls_final_small", _p(status), _p(n_acc), _p(ds), _p(reg), _p(stats), _p(mad), _p(target),
                         _p(logsig), _p(beta), _p(gamma), _p(the_m), _p(sig), _p(pack), _p(sm), d, P, _p(small),
                         self._stream())
        self.launches += 1

    # ---- kernel 3, fused: three launches with their reductions, exchanges and solves inside (csrc/tail.cu) --------
    def tail_supported(self, d, P) -> bool:
        return bool(self.lib.abcdls_tail_supported(int(d), int(P)))

    def tail_workspace(self, d, P) -> torch.Tensor:
        key = f"tail{d}x{P}"
        t = self._ws.get(key)
        if t is None:
            nbytes = int(self.lib.abcdls_tail_workspace_bytes(self.handle.h, int(d), int(P)))
            t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)     # the ticket counters start at zero
            self._ws[key] = t
        return t

    def tail_gather_fit(self, peer, cand_ss, slot, perm, ss, par, dist_acc, idx, n_acc, row_offset, cap, d, P, mad,
                        target, ds, loclinear, xs, y, ss_out, w, ws, status):
        """par: (P, n) view - unit stride along rows (column-major table) or strides (1, pitch) (row-major table)."""
        cm = par.stride(1) == 1
        self.handle.call("abcdls_tail_gather_fit", peer, _p(cand_ss), cand_ss.stride(0) if cand_ss is not None else 0,
                         _p(slot), _p(perm), _p(ss), ss.stride(0) if ss is not None else 0,
                         _p(par) if cm else None, par.stride(0) if cm else 0,
                         None if cm else _p(par), 0 if cm else par.stride(1),
                         _p(dist_acc), _p(idx), _p(n_acc), row_offset, cap, d, P, _p(mad), _p(target), _p(ds),
                         int(loclinear), _p(xs), _p(y), _p(ss_out), _p(w), _p(ws), ws.numel(), _p(status), self._stream())
        self.launches += 1

    def copy_out(self, segs, memo=None):
        """segs: up to four (tensor2d_or_1d, rows, cols) with unit stride along cols -> ONE dense float64 block holding
        the segments back to back (int64 data travels as raw 8-byte words), in one launch; returns the block.
        memo (a dict that lives as long as the source buffers, e.g. the graph plan): keeps the marshalled arguments."""
        shape = tuple((r, c) for _, r, c in segs)
        co = memo.get("copy_out") if memo is not None else None
        if co is None or co[0] != shape:
            n = len(segs)
            src = (C.c_void_p * n)(*[t.data_ptr() for t, _, _ in segs])
            ld = (C.c_int64 * n)(*[(t.stride(0) if t.dim() == 2 else max(c, 1)) for t, _, c in segs])
            rows = (C.c_int64 * n)(*[r for _, r, _ in segs])
            cols = (C.c_int64 * n)(*[c for _, _, c in segs])
            co = (shape, n, src, ld, rows, cols, sum(r * c for r, c in shape))
            if memo is not None:
                memo["copy_out"] = co
        _, n, src, ld, rows, cols, total = co
        out = torch.empty(total, dtype=torch.float64, device=self.device)
        self.handle.call("abcdls_copy_out", n, src, ld, rows, cols, _p(out), self._stream())
        self.launches += 1
        return out

    def abc_single(self, target, param, sumstat, tol, method, hcorr=True):
        """The call-level C entry point (csrc/single.cu: abc() on one GPU as ONE C call, one-pass path only) - what a
        non-Python host binds instead of mirroring engine.py.  Returns the raw output blocks (leading dimension cap)
        and the packed small buffer; nothing is synchronised.  The engine itself keeps sequencing the stages (it needs
        the exchanges of several ranks, the fallback ladder and CUDA graphs in between)."""
        d, n = sumstat.shape
        P = param.shape[0]
        lin = method == "loclinear"
        nacc = int(self.lib.abcdls_abc_single_nacc(n, float(tol)))
        cap = min(nacc, n)
        ws = self.workspace("single", int(self.lib.abcdls_abc_single_workspace_bytes(self.handle.h, n, d, P, float(tol))))
        idx, dacc = self.new(cap, dtype=torch.int64), self.new(cap)
        y, ss_out = self.new(P, cap), self.new(d, cap)
        xs = w = adj = res = None
        if lin:
            xs, w, adj, res = self.new(d, cap), self.new(cap), self.new(P, cap), self.new(P, cap)
        small = self.new(self.final_small_len(d, P, lin))
        cm = param.stride(1) == 1
        self.handle.call("abcdls_abc_single", _p(target), _p(param) if cm else None, param.stride(0) if cm else 0,
                         None if cm else _p(param), 0 if cm else param.stride(1), _p(sumstat), sumstat.stride(0), n, d, P,
                         float(tol), int(lin), int(bool(hcorr)), _p(idx), _p(dacc), _p(y), _p(ss_out), _p(xs), _p(w),
                         _p(adj), _p(res), _p(small), _p(ws), ws.numel(), self._stream())
        return dict(nacc=nacc, cap=cap, idx=idx, dist=dacc, y=y, ss_out=ss_out, xs=xs, w=w, adj=adj, res=res, small=small)

    def tail_resid_fit(self, peer, xs, y, w, n_acc, cap, d, P, ws, status):
        self.handle.call("abcdls_tail_resid_fit", peer, _p(xs), _p(y), _p(w), _p(n_acc), cap, d, P, _p(ws),
                         ws.numel(), _p(status), self._stream())
        self.launches += 1

    def tail_adjust_final(self, peer, xs, y, w, n_acc, ds, cap, d, P, loclinear, hcorr, mad, target, adj, res, ws, small,
                          status):
        self.handle.call("abcdls_tail_adjust_final", peer, _p(xs), _p(y), _p(w), _p(n_acc), _p(ds), cap, d, P,
                         int(loclinear), int(hcorr), _p(mad), _p(target), _p(adj), _p(res), _p(ws), ws.numel(), _p(small),
                         _p(status), self._stream())
        self.launches += 1

    # ---- per-column mode as one pass (csrc/columns.cuh) ---------------------------------------------------------------
    def columns_supported(self, ss, n, P, ld) -> bool:
        return bool(self.lib.abcdls_columns_supported(_p(ss), n, P, ld))

    def columns_front(self, ss, n, n_global, P, ld, target, nacc, hint, med, mad, status, comm=None):
        """sample -> pass -> medians / MADs -> exact distances of the T-band members.
        -> (ws, dist table (P, capT) view, cnt (P,) int64 device tensor)"""
        world = comm.world if comm is not None else 1
        ws = self.workspace(f"columns{P}", self.lib.abcdls_columns_workspace_bytes(self.handle.h, n, P, hint))
        p0, c0, p1, c1, p2 = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p()
        st = self._stream()
        self.handle.call("abcdls_columns_sample", _p(ss), n, P, ld, _p(target), nacc, n_global, hint, world, _p(ws),
                         ws.numel(), C.byref(p0), C.byref(c0), st)
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.float64, 8))
        with self._timed("k_columns_pass"):
            self.handle.call("abcdls_columns_pass", _p(ss), n, P, ld, _p(target), hint, world, _p(ws), ws.numel(),
                             C.byref(p0), C.byref(c0), C.byref(p2), st)
        if world > 1:
            comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
            comm.allreduce_sum(self._view(ws, p2, C.c_size_t(P), torch.int64, 8))
        gs = gc = None
        with self._timed("columns_refine"):
            for phase in range(4):
                self.handle.call("abcdls_columns_refine", n, n_global, P, hint, phase, world, _p(gs), _p(gc), _p(ws),
                                 ws.numel(), _p(med), _p(mad), _p(status), C.byref(p0), C.byref(c0), C.byref(p1),
                                 C.byref(c1), st)
                if world > 1:
                    if phase in (0, 2):
                        gs, gc = comm.allgather_ragged(self._view(ws, p0, c0, torch.float64, 8),
                                                       self._view(ws, p1, c1, torch.int32, 4))
                    elif phase == 1:
                        comm.allreduce_sum(self._view(ws, p0, c0, torch.int64, 8))
        cnt = self.new(P, dtype=torch.int64)
        cap = C.c_int64()
        self.handle.call("abcdls_columns_exact", n, P, hint, _p(mad), _p(target), _p(ws), ws.numel(), C.byref(p0),
                         C.byref(cap), _p(cnt), _p(status), st)
        dist = self._view(ws, p0, C.c_size_t(P * cap.value), torch.float64, 8).view(P, cap.value)
        self.launches += 22
        return ws, dist, cnt

    def columns_count(self, ws, n, P, hint, nacc, sel, mad, target, cnt, status) -> torch.Tensor:
        p0 = C.c_void_p()
        self.handle.call("abcdls_columns_count", n, P, hint, nacc, _p(sel), _p(mad), _p(target), _p(cnt), _p(ws),
                         ws.numel(), C.byref(p0), _p(status), self._stream())
        self.launches += 1
        return self._view(ws, p0, C.c_size_t(2 * P), torch.int64, 8).view(P, 2)

    def columns_rowkeys(self, ws, n, P, hint, sel, cnt, cap) -> torch.Tensor:
        p0 = C.c_void_p()
        self.handle.call("abcdls_columns_rowkeys", n, P, hint, _p(sel), _p(cnt), _p(ws), ws.numel(), C.byref(p0),
                         self._stream())
        self.launches += 1
        return self._view(ws, p0, C.c_size_t(P * cap), torch.float64, 8).view(P, cap)

    def columns_accept(self, peer, ws, n, P, hint, sel, med, mad, cnt, thr, par, small, status, lists=None):
        acc_cap, row, val, dst, acnt = (0, None, None, None, None) if lists is None else lists
        self.handle.call("abcdls_columns_accept", peer, n, P, hint, _p(sel), _p(med), _p(mad), _p(cnt), _p(thr),
                         _p(par), par.stride(0), par.stride(1), acc_cap, _p(row), _p(val), _p(dst), _p(acnt), _p(ws),
                         ws.numel(), _p(small), _p(status), self._stream())
        self.launches += 2

    # ---- summary.abc rows beyond reductions (csrc/summary.cu) ----------------------------------------------------------
    def wselect(self, x, w, n, P, taus) -> torch.Tensor:
        """(P, ntau) weighted order statistics of the columns of x (P, n) column-major."""
        ntau = taus.numel()
        ws = self.workspace("wselect", self.lib.abcdls_wselect_workspace_bytes(P * ntau))
        out = self.new(P, ntau)
        self.handle.call("abcdls_wselect", _p(x), x.stride(0) if P > 1 else n, _p(w), n, P, _p(taus), ntau, _p(out), _p(ws), ws.numel(),
                         self._stream())
        self.launches += 13
        return out

    def col_moments(self, x, n, P) -> torch.Tensor:
        out = self.new(P, 2)
        self.handle.call("abcdls_col_moments", _p(x), x.stride(0) if P > 1 else n, n, P, _p(out), self._stream())
        self.launches += 1
        return out

    def density_mode(self, x, w, n, P, lo, up, frm, to, bw, wnorm) -> torch.Tensor:
        ws = self.workspace("density", self.lib.abcdls_density_workspace_bytes(P))
        out = self.new(P)
        self.handle.call("abcdls_density_mode", _p(x), x.stride(0) if P > 1 else n, _p(w), n, P, _p(lo), _p(up), _p(frm), _p(to), _p(bw),
                         _p(wnorm), _p(out), _p(ws), ws.numel(), self._stream())
        self.launches += 2
        return out

    def small_supported(self, n, nacc) -> bool:
        return bool(self.lib.abcdls_small_supported(int(n), int(nacc)))

    def small_columns(self, ss, par, n, col, hold, target, nacc, loclinear, hcorr, out):
        """B one-CTA abc() problems with d = P = 1 in ONE launch; ss (cols, n) column-major, par (cols, n) any strides."""
        self.handle.call("abcdls_small_columns", _p(ss), ss.stride(0), _p(par), par.stride(0), par.stride(1), n, _p(col),
                         _p(hold), _p(target), col.numel(), nacc, int(loclinear), int(hcorr), _p(out), self._stream())
        self.launches += 1

    def postpr_counts(self, model_id, idx, n_acc, row_offset, cap, K, counts):
        self.handle.call("abcdls_postpr_counts", _p(model_id), _p(idx), _p(n_acc), row_offset, cap, K, _p(counts),
                         self._stream())
        self.launches += 2

    def postpr_final(self, peer, status, n_acc, ds, mad, d, counts, K, small):
        self.handle.call("abcdls_postpr_final", peer, _p(status), _p(n_acc), _p(ds), _p(mad), d, _p(counts), K, _p(small),
                         self._stream())
        self.launches += 1

    # ---- kernel 4 -------------------------------------------------------------------------------
    def smc_workspace(self, n, P):
        return self.workspace("smc", self.lib.abcdls_smc_workspace_bytes(n, P))

    def smc_oldrange(self, table, n, P, ld, minmax):
        ws = self.smc_workspace(n, P)
        self.handle.call("abcdls_smc_oldrange", _p(table), n, P, ld, _p(minmax), _p(ws), ws.numel(), self._stream())
        self.launches += 3

    def smc_box_filter(self, table, n, P, ld, lo, hi, row_offset, cap, idx, n_out, status):
        ws = self.smc_workspace(n, P)
        self.handle.call("abcdls_smc_box_filter", _p(table), n, P, ld, _p(lo), _p(hi), row_offset, cap, _p(idx),
                         _p(n_out), _p(ws), ws.numel(), _p(status), self._stream())
        self.launches += 4
