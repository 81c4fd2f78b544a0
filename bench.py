#!/usr/bin/env python
"""Benchmark of the ABC acceptance-and-adjustment hot path (BASELINE.json metric:
"simulations scored/sec (ABC loclinear)").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of R abc's algorithm

A "step" is one full abc(method="loclinear") call (MAD scaling -> distance -> tolerance cut -> local-linear
adjustment with hcorr -> summary reductions) over one synthetic table (SURVEY.md §8d recipe).
Default workload: BASELINE config 5 headline shape, N = 1e8 simulations x 16 summaries x 16 parameters, tol 0.01,
sharded by rows over the ranks (strong scaling: the table is fixed, N_local = N / n_gpus).
Other BASELINE configs (one JSON line each, same contract keys; `--impl reference` works for all of them):
    --config cfg3            abc loclinear joint, N = 1e7 x 10 x 10, tol 0.001
    --workload postpr        cfg2: postpr rejection, N = 3e6, K = 3 models, tol 0.05 (68 B / simulation)
    --workload smc           cfg4: 5 SMC rounds x 1e7 simulations, 12 parameters, per-column rejection -> min / max ->
                             range rules -> box filter (tables regenerated inside the new box between rounds, untimed)
    --workload sweep         cfg5 at N = 1e6, 3e6, 1e7, 3e7, 1e8 ("sweep": one entry per N)
    --workload cv4abc        SURVEY 8f-1 at the cfg1 shape

`value`  : whole-job sims/s, tables resident in HBM, CUDA events, max over ranks.
`e2e`    : same metric through the host-buffer API (pinned host tables -> H2D -> abc -> D2H of the
           posterior sample and summary) per step.
`roofline`: dominant table-streaming kernel, algorithmic bytes / its measured duration vs MEASURED_PEAKS.json.
`cpu_baseline`: the numpy oracle (port of CRAN abc 2.2.1's algorithm; R itself is absent on the box) timed on
           the host, 1 core (R abc is single threaded), on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simulations scored/sec (ABC loclinear)"
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture at the default
# workload (k_fused_pass: profiles/r02_ncu_full_pass_and_fine_sample.txt; k_bracket_pass: profiles/r01_*); None until a capture of the current kernel is committed
TRAFFIC = {"k_fused_pass": 12.814157e9 + 0.376328e9, "k_bracket_pass": 12.800307e9 + 0.494554e9, "k_dist_pass": None}
UNIT = "sims/s"
PRESETS = {"cfg5": dict(rows=1e8, stats=16, params=16, tol=0.01),
           "cfg3": dict(rows=1e7, stats=10, params=10, tol=0.001)}


def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg5", choices=list(PRESETS),
                    help="BASELINE.json config whose shape fills --rows / --stats / --params / --tol when not given")
    ap.add_argument("--rows", type=float, default=None, help="global simulations N")
    ap.add_argument("--stats", type=int, default=None)
    ap.add_argument("--params", type=int, default=None)
    ap.add_argument("--tol", type=float, default=None)
    ap.add_argument("--method", default="loclinear")
    ap.add_argument("--cpu-rows", type=float, default=None,
                    help="bounded sample for the CPU arm (default: min(N, 1e7) rows, SURVEY 8d)")
    ap.add_argument("--param-layout", default="row", choices=["row", "col"],
                    help="memory layout of the parameter table: row = n x P C order (numpy / pandas, what ABC-DLS holds), "
                         "col = P x n (R's as.matrix); sumstat is always column-major")
    ap.add_argument("--workload", default="abc", choices=["abc", "postpr", "smc", "sweep", "cv4abc"],
                    help="abc: the headline call (default).  postpr / smc / sweep: BASELINE configs 2 / 4 / 5-sweep.  "
                         "cv4abc: SURVEY 8f-1, leave-one-out cross-validation at the cfg1 shape (10 000 test rows, 19 "
                         "parameters, per-column, nval 100), launch-bound")
    ap.add_argument("--rounds", type=int, default=5, help="SMC rounds of --workload smc")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-f32", action="store_true", help="skip the extra end-to-end leg with float32 host tables")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--clock-interval", type=float, default=0.02, help="seconds between NVML clock samples")
    a = ap.parse_args()
    preset = dict(PRESETS[a.config])
    if a.workload == "postpr":
        preset = dict(rows=3e6, stats=3, params=3, tol=0.05)
    elif a.workload == "smc":
        preset = dict(rows=1e7, stats=12, params=12, tol=0.01)
    for k_, v_ in preset.items():
        if getattr(a, k_) is None:
            setattr(a, k_, v_)
    if a.cpu_rows is None:
        a.cpu_rows = min(a.rows, 1e7)
    return a


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML in-process (forking
    nvidia-smi from a process that maps tens of GB stalls the launching thread for milliseconds)."""

    def __init__(self, index: int, interval: float = 0.02):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.interval, self.call_ms = interval, []
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.max_mhz, self.nv, self.h = None, None, None
        try:  # import + nvmlInit are slow (tens of ms, GIL held): do them before the timed region
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            # the first query of each kind takes ~15 ms inside the driver and stalls kernel launches meanwhile
            nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            self.nv = nv
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv, h = self.nv, self.h
        if nv is None:
            return
        while not self.stop.is_set():
            try:
                t0 = time.perf_counter()
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
                self.call_ms.append((time.perf_counter() - t0) * 1e3)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            self.stop.wait(self.interval)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=5)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled: " + getattr(self, "err", "")]}
        sm = sorted(r[0] for r in self.rows)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "hw_power_brake_slowdown": 0x80}
        reasons = [n for n, b in bits.items() if any(r[1] & b for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "nvml_ms_per_sample": round(sum(self.call_ms) / len(self.call_ms), 3)}


def probe_r():
    """Is the real reference path (rpy2 -> R package abc) usable on this box?  Probed at run time (SURVEY 8c-4,
    BASELINE.md §4 step 1): never assumed."""
    import shutil
    import subprocess
    info = {"Rscript": None, "r_abc": False, "rpy2": False}
    exe = shutil.which("Rscript")
    info["Rscript"] = exe
    if exe:
        try:
            out = subprocess.run([exe, "-e", "suppressMessages(library(abc)); cat(as.character(packageVersion('abc')))"],
                                 capture_output=True, text=True, timeout=120)
            info["r_abc"] = out.returncode == 0
            info["r_abc_version"] = out.stdout.strip()[-20:] if out.returncode == 0 else out.stderr.strip()[-200:]
        except Exception as e:  # pragma: no cover
            info["r_abc_version"] = repr(e)
    try:
        import rpy2  # noqa: F401
        info["rpy2"] = True
    except Exception as e:
        info["rpy2_error"] = type(e).__name__
    return info


R_SCRIPT = r"""
suppressMessages(library(abc))
a <- commandArgs(trailingOnly = TRUE)
dir <- a[1]; n <- as.integer(a[2]); d <- as.integer(a[3]); P <- as.integer(a[4]); tol <- as.numeric(a[5])
method <- a[6]; steps <- as.integer(a[7]); kind <- a[8]
rd <- function(f, k) matrix(readBin(file.path(dir, f), "double", n * k, size = 8), nrow = n, ncol = k)
ss <- rd("ss.bin", d); tg <- readBin(file.path(dir, "tgt.bin"), "double", d, size = 8)
if (kind == "postpr") {
  idx <- as.character(readBin(file.path(dir, "idx.bin"), "integer", n, size = 4))
  t0 <- proc.time()[["elapsed"]]
  for (i in seq_len(steps)) r <- postpr(target = tg, index = idx, sumstat = ss, tol = tol, method = "rejection")
  dt <- (proc.time()[["elapsed"]] - t0) / steps
  writeBin(as.double(table(r$values)), file.path(dir, "counts.out"), size = 8)
} else {
  par <- rd("par.bin", P)
  t0 <- proc.time()[["elapsed"]]
  for (i in seq_len(steps)) { r <- abc(target = tg, param = par, sumstat = ss, tol = tol, method = method); s <- summary(r, print = FALSE) }
  dt <- (proc.time()[["elapsed"]] - t0) / steps
  writeBin(as.integer(which(r$region)), file.path(dir, "region.out"), size = 4)
  v <- if (method == "rejection") r$unadj.values else r$adj.values
  writeBin(as.double(v), file.path(dir, "values.out"), size = 8)
}
cat(sprintf("SECONDS_PER_CALL %.9f\n", dt))
"""


def r_arm(kind, tables, tol, method, steps):
    """Times the REAL reference arithmetic (R package abc through Rscript; same inputs, written as raw doubles) and
    keeps its outputs under baseline/_ref/ as golden vectors for the oracle.  Returns (seconds per call,
    dict of golden arrays) or raises."""
    import subprocess
    import tempfile
    import numpy as np
    gold = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(gold, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        n, d = tables["ss"].shape
        P = tables["par"].shape[1] if "par" in tables else 0
        np.asfortranarray(tables["ss"]).T.tofile(os.path.join(td, "ss.bin"))          # column-major doubles
        tables["tgt"].astype(np.float64).tofile(os.path.join(td, "tgt.bin"))
        if "par" in tables:
            np.asfortranarray(tables["par"]).T.tofile(os.path.join(td, "par.bin"))
        if "idx" in tables:
            tables["idx"].astype(np.int32).tofile(os.path.join(td, "idx.bin"))
        script = os.path.join(td, "arm.R")
        open(script, "w").write(R_SCRIPT)
        out = subprocess.run(["Rscript", script, td, str(n), str(d), str(P), repr(float(tol)), method, str(steps), kind],
                             capture_output=True, text=True, timeout=3000)
        if out.returncode != 0:
            raise RuntimeError("Rscript failed: " + out.stderr[-500:])
        dt = float([l for l in out.stdout.splitlines() if l.startswith("SECONDS_PER_CALL")][-1].split()[1])
        res = {}
        for f, dt_ in (("region.out", np.int32), ("values.out", np.float64), ("counts.out", np.float64)):
            if os.path.exists(os.path.join(td, f)):
                res[f] = np.fromfile(os.path.join(td, f), dtype=dt_)
                res[f].tofile(os.path.join(gold, f"{kind}_{n}x{d}_{f}"))
        return dt, res


def cpu_tables(args, kind):
    """The bounded sample of the workload the CPU arm runs on (row-prefix of the same seeded tables)."""
    from abc_dls_b200 import synth
    n = int(min(args.cpu_rows, args.rows))
    if kind == "postpr":
        ids, ss, tgt = synth.postpr_tables(n, "cpu", K=args.stats)
        return {"ss": ss.numpy().T.copy(order="F"), "tgt": tgt.numpy(), "idx": ids.numpy()}
    d = args.stats if kind == "abc" else args.params
    par, ss, tgt = synth.abc_tables(n, d, args.params, "cpu")
    return {"par": par.numpy().T.copy(order="F"), "ss": ss.numpy().T.copy(order="F"), "tgt": tgt.numpy()}


def cpu_arm(args, steps, warmup, kind="abc"):
    """The reference's CPU implementation of the path, timed on the host on a bounded sample of the same workload:
    real R (Rscript + package abc) when the box has it, else the numpy oracle (a port of CRAN abc 2.2.1's algorithm).
    R abc is single threaded, so is the port: cores = 1.  value = rows timed / seconds per call."""
    import numpy as np
    import torch
    from oracle import abc_oracle as O
    torch.set_num_threads(1)
    T = cpu_tables(args, kind)
    n = T["ss"].shape[0]
    probe = probe_r()
    method = "rejection" if kind in ("postpr", "smc") else args.method
    what = {"abc": f"abc({method}) joint {n} rows x {args.stats} stats x {args.params} params, tol {args.tol}",
            "postpr": f"postpr(rejection) {n} rows x K = {args.stats} models, tol {args.tol}",
            "smc": f"one SMC round: {args.params} per-column abc(rejection) on {n} rows, tol {args.tol}, + range rules + "
                   f"box filter"}[kind]
    if probe["r_abc"] and kind != "smc":
        try:
            dt, gold = r_arm(kind, T, args.tol, method, max(1, steps))
            note = ""
            if "region.out" in gold:   # diff the oracle against the real thing while we are here
                o = O.abc(T["tgt"], T["par"], T["ss"], args.tol, method)
                same = np.array_equal(gold["region.out"].astype(np.int64) - 1, o.idx)
                note = f"; accepted set of the numpy oracle identical to R's: {same}"
            return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "reference", "rows_timed": n, "r_probe": probe,
                    "sample": f"R package abc {probe.get('r_abc_version', '')} through Rscript, {what}, {steps} calls, "
                              f"{dt * 1e3:.1f} ms/call (single threaded: 1 of {os.cpu_count()} cores){note}"}, dt
        except Exception as e:  # pragma: no cover  (no R in the build container)
            probe["r_arm_error"] = repr(e)[:300]

    def one():
        if kind == "postpr":
            return O.postpr(T["tgt"], T["idx"], T["ss"], args.tol)
        if kind == "smc":
            return O.smc_round(T["tgt"], T["par"], T["ss"], args.tol, "rejection", T["par"], 0.95, 0.01,
                               T["par"].min(axis=0) - 1.0, T["par"].max(axis=0) + 1.0)
        r = O.abc(T["tgt"], T["par"], T["ss"], args.tol, method)
        v = r.adj_values if r.adj_values is not None else r.unadj_values      # what the timed GPU call also returns:
        return v.min(axis=0), v.max(axis=0), (v * r.weights[:, None]).sum(axis=0) / r.weights.sum()   # min / max / mean

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    absent = ", ".join(k_ for k_, ok in (("Rscript", probe["Rscript"]), ("R package abc", probe["r_abc"]),
                                         ("rpy2", probe["rpy2"])) if not ok)
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port", "rows_timed": n, "r_probe": probe,
            "sample": f"numpy oracle (port of CRAN abc 2.2.1; probed at run time, absent on this box: {absent or 'nothing'}), "
                      f"{what}, {steps} calls, {dt * 1e3:.1f} ms/call; R abc is single threaded so 1 of {os.cpu_count()} "
                      f"cores is used; value = rows timed / time, i.e. per-simulation throughput at {n} rows"}, dt


def cv_workload(args):
    """cv4abc as ABC-DLS runs it before every parameter estimate (ABC.py:1896-1904): per parameter column, nval = 100
    held-out rows, loclinear, tol 0.01, on the test-set sized table (cfg1: 10 000 rows x 19 parameters).  GPU: every fit is
    one thread block of csrc/small.cu; timed as one launch for all columns, as one call per column, and the kernel alone
    (CUDA events).  CPU: the oracle on a bounded sample of the same fits."""
    import numpy as np
    import torch
    from abc_dls_b200 import synth, cv
    from abc_dls_b200.engine import AbcEngine
    from oracle import abc_oracle as O
    N, P, nval, tol = 10000, 19, 100, 0.01
    dev = torch.device("cuda", 0)
    eng = AbcEngine(dev)
    par, ss, _ = synth.abc_tables(N, P, P, dev)
    samp = np.random.default_rng(1).choice(N, nval, replace=False)

    def run_calls():     # as ABC-DLS calls it: one cv4abc per parameter column (P launches, P blocking reads)
        return [cv.cv4abc(eng, par[c:c + 1], ss[c:c + 1], nval, tol, "loclinear", cvsamples=samp) for c in range(P)]

    def run():           # the whole "Cv error independently" loop as ONE launch of P x nval thread blocks
        return cv.cv4abc_columns(eng, par, ss, nval, tol, "loclinear", cvsamples=samp)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        l0 = eng.k.launches
        t0 = time.perf_counter()
        for _ in range(reps):
            o_ = fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / reps, (eng.k.launches - l0) // reps, o_

    dt_calls, launches_calls, _ = timed(run_calls, 5)
    dt, launches, out = timed(run, 20)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    col_t = torch.arange(P, dtype=torch.int32, device=dev).repeat_interleave(nval)
    hold_t = torch.from_numpy(np.tile(samp, P)).to(dev)
    ko = eng.k.new(P * nval, 8)
    import math
    nacc = int(math.ceil((N - 1) * tol))
    ev0.record()
    for _ in range(20):
        eng.k.small_columns(ss, par, N, col_t, hold_t, None, nacc, True, True, ko)
    ev1.record()
    torch.cuda.synchronize(dev)
    kernel_ms = ev0.elapsed_time(ev1) / 20
    cpu_nval = 10
    pn, sn = par.cpu().numpy().T, ss.cpu().numpy().T
    t0 = time.perf_counter()
    ref = [O.cv4abc(pn[:, c], sn[:, c], cpu_nval, tol, "loclinear", cvsamples=samp[:cpu_nval]) for c in range(P)]
    dt_cpu = time.perf_counter() - t0
    err = max(float(np.max(np.abs(out[c].estim[tol][:cpu_nval] - ref[c].estim[tol]) / np.abs(ref[c].true).max()))
              for c in range(P))
    fits = P * nval
    print(json.dumps({"metric": "leave-one-out abc() fits/sec (cv4abc, per column, loclinear)", "value": fits / dt,
                      "unit": "fits/s", "n_gpus": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "dtype": "f64",
                      "data": "synthetic", "config": {"workload": f"cfg1 cv4abc: N={N} test rows, {P} parameter columns "
                                                                  f"(d=1 each), nval={nval}, tol={tol}, loclinear, median"},
                      "bound": "shared-memory radix selects inside one CTA per fit (csrc/small.cu); the streaming path "
                               "needs ~70 launches per fit at this size and measured 2 031 fits/s",
                      "gpu_launches": launches, "kernel_ms": kernel_ms, "kernel_fits_per_s": fits / (kernel_ms * 1e-3),
                      "per_column_calls": {"value": fits / dt_calls, "ms": dt_calls * 1e3, "gpu_launches": launches_calls},
                      "max_rel_err_vs_oracle": err,
                      "cpu_baseline": {"value": P * cpu_nval / dt_cpu, "unit": "fits/s", "cores": 1, "kind": "port",
                                       "sample": f"numpy oracle cv4abc, same table, {P} columns x {cpu_nval} held-out rows"}}))


def bind_to_gpu_numa(index: int):
    """Run this process (and first-touch its pinned host buffers) on the CPUs NVML reports as local to GPU `index`: the
    eight concurrent H2D copies of an 8-rank end-to-end step then read host memory of their own socket instead of
    contending for one.  Host-side placement only; returns what was done for the JSON line."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        mask = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(local & allowed)
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            return {"bound": True, "cpus": len(use), "of": len(allowed)}
        return {"bound": False, "cpus": len(use), "of": len(allowed)}
    except Exception as e:  # pragma: no cover
        return {"bound": False, "error": type(e).__name__}


class Ctx:
    """One process per GPU: rank / world from the launcher, the engine, barrier + device sync, max over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from abc_dls_b200.engine import AbcEngine
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.numa = bind_to_gpu_numa(self.local_rank) if self.world > 1 else {"bound": False}
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.eng = AbcEngine(self.dev)

    def shard(self, N):
        per = (N + self.world - 1) // self.world
        r0, r1 = min(self.rank * per, N), min((self.rank + 1) * per, N)
        return r0, r1 - r0

    def sync(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def allmax(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def digest(self, idx, values):
        """Order-independent fingerprints over ALL ranks: of the accepted global row numbers (wrapping 64-bit sum and
        xor) and of the accepted parameter rows (wrapping sum of their bit patterns) - bit-identical at 1 / 2 / 4 / 8 GPUs
        iff the accepted set is; plus the plain sum of the adjusted values (reduction order differs across GPU counts:
        compare to ~1e-12)."""
        torch = self.torch
        i64 = idx.to(torch.int64)
        parts = torch.stack([i64.sum(), values.contiguous().view(torch.int64).sum()])
        x = i64.new_zeros(1)
        if i64.numel():
            x = torch.tensor([0], dtype=torch.int64, device=self.dev)
            h = i64.clone()
            while h.numel() > 1:                      # xor tree on the device
                if h.numel() & 1:
                    h = torch.cat([h, h.new_zeros(1)])
                h = h[0::2] ^ h[1::2]
            x = h
        if self.world > 1:
            self.dist.all_reduce(parts, op=self.dist.ReduceOp.SUM)
            xs = [torch.zeros_like(x) for _ in range(self.world)]
            self.dist.all_gather(xs, x)
            x = xs[0]
            for t in xs[1:]:
                x = x ^ t
        return {"rows_sum_u64": f"{int(parts[0].item()) & (2 ** 64 - 1):016x}", "rows_xor": f"{int(x.item()) & (2 ** 64 - 1):016x}",
                "unadj_bits_sum_u64": f"{int(parts[1].item()) & (2 ** 64 - 1):016x}"}

    def allsum(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_region(ctx, args, step):
    """W >= 3 untimed steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream,
    max over ranks; SM clocks / throttle reasons sampled meanwhile."""
    torch = ctx.torch
    for _ in range(max(args.warmup, 3)):   # call 1 runs eagerly, call 2 captures the CUDA graph, call 3+ replay it
        step()
    ctx.sync()
    l0 = ctx.eng.k.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(ctx.local_rank, args.clock_interval)
    res = None
    with sampler as clocks:
        ctx.sync()
        if hasattr(ctx.eng, "host_profile"):
            ctx.eng.host_profile = {k_: 0 for k_ in ctx.eng.host_profile}
            ctx.eng._t_last_return = None
        ev0.record()
        for _ in range(args.steps):
            res = step()
        ev1.record()
        ctx.sync()
        if hasattr(ctx.eng, "host_profile"):
            ctx.host_profile_timed = dict(ctx.eng.host_profile)
    ms = ctx.allmax(ev0.elapsed_time(ev1)) / args.steps
    return ms, ctx.eng.k.launches - l0, clocks.summary(), res


def stage_times(ctx, args, step):
    """per-stage / per-kernel durations: the same call enqueued eagerly with CUDA events around every stage"""
    eng = ctx.eng
    eng.timers = {}
    reps = min(args.steps, 5)
    for _ in range(reps):
        step()
    ctx.sync()
    tms = eng.timer_ms()
    eng.timers = None
    return {name: {"launches": len(v), "avg_ms": sum(v) / len(v), "total_ms_per_step": sum(v) / reps}
            for name, v in tms.items()}


def streaming_roofline(kern, d, n_local, bytes_per_sim, N, ms_step, world):
    """Roofline of the dominant table-streaming kernel (per launch, this rank): CUDA events on the launching stream
    around the C-ABI call that enqueues it; plus the whole call against SURVEY 8d's bytes-per-simulation model."""
    peak, peak_src = _peaks()
    alg = {"k_fused_pass": 8.0 * d * n_local,        # THE one read of the N x d table (medians, MADs and the cut)
           "k_columns_pass": 8.0 * d * n_local,      # per-column mode: one read of the N x P statistics table
           "k_bracket_pass": 8.0 * d * n_local,      # one read of the N x d table (all medians + MADs)
           "k_dist_pass": 8.0 * d * n_local,         # one read of the table; dist is not materialised
           "scale_dist": (8.0 * d + 8.0) * n_local}
    for name in list(kern):   # radix-select passes stream the table only on the exact path (names carry the job shape)
        if name.startswith("select_hist") and name.endswith(f"x{n_local}]"):
            alg[name] = 8.0 * float(name.split("[")[1].split("x")[0]) * n_local
    streaming = [k_ for k_ in kern if k_ in alg]
    whole = {"algorithmic_bytes_per_sim": bytes_per_sim,
             "achieved_GBps": bytes_per_sim * N / (ms_step * 1e-3) / 1e9 / world,
             "frac_of_peak_per_gpu": bytes_per_sim * N / (ms_step * 1e-3) / 1e9 / world / peak}
    if not streaming:
        return {"bound": "hbm", "kernel": None, "achieved": whole["achieved_GBps"], "peak": peak, "unit": "GB/s",
                "frac": whole["frac_of_peak_per_gpu"], "traffic": None, "peak_source": peak_src, "whole_call": whole,
                "stages_ms": {k_: round(v_["avg_ms"], 4) for k_, v_ in kern.items()}}
    dom = max(streaming, key=lambda k_: kern[k_]["total_ms_per_step"])
    achieved = alg[dom] / (kern[dom]["avg_ms"] * 1e-3) / 1e9
    per_kernel = {k_: {"algorithmic_bytes": alg[k_], "avg_ms": kern[k_]["avg_ms"],
                       "GBps": alg[k_] / (kern[k_]["avg_ms"] * 1e-3) / 1e9,
                       "frac": alg[k_] / (kern[k_]["avg_ms"] * 1e-3) / 1e9 / peak} for k_ in streaming}
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": TRAFFIC.get(dom) if n_local == 100_000_000 and d == 16 else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": kern[dom]["avg_ms"],
            "streaming_kernels": per_kernel, "whole_call": whole,
            "stages_ms": {k_: round(v_["avg_ms"], 4) for k_, v_ in kern.items()}}


def engine_facts(ctx):
    eng = ctx.eng
    host_ms = eng.host_enqueue_ms[-8:]
    hp = dict(getattr(ctx, "host_profile_timed", None) or eng.host_profile)   # of the timed calls
    nc = max(1, hp.pop("calls"))
    return {"fast_path_misses": eng.fast_misses, "host_enqueue_ms": round(sum(host_ms) / max(1, len(host_ms)), 3),
            "host_profile_us": {k_: round(v_ / nc, 1) for k_, v_ in hp.items()},
            "cuda_graph": bool(eng._graphs), "front_half": dict(eng.path_calls),
            "collectives": {"total": eng.comm.calls, "nvlink_mailbox_kernels": eng.comm.peer_calls,
                            "fused_into_compute_kernels": eng.comm.fused_calls,
                            "nccl": eng.comm.calls - eng.comm.peer_calls - eng.comm.fused_calls}}


def abc_config(args, world, N=None):
    N = int(args.rows if N is None else N)
    name = {(100_000_000, 16, 16): "cfg5", (10_000_000, 10, 10): "cfg3"}.get((N, args.stats, args.params), "custom")
    return {"workload": f"{name} abc(method={args.method}, hcorr) joint: N={N:d} sims x d={args.stats} "
                        f"stats x P={args.params} params, tol={args.tol}, fp64 tables: sumstat column-major, param "
                        f"{'row-major (n x P C order, as numpy / pandas hand the frame over)' if args.param_layout == 'row' else 'column-major'}",
            "param_layout": args.param_layout, "rows": N, "stats": args.stats, "params": args.params, "tol": args.tol,
            "sharding": f"rows/{world}", "l2": "tables larger than L2 (no flush needed)" if N * args.stats * 8 > 3e8
            else "L2 flushed by the call itself: every call streams > 126 MB of candidate / scratch buffers is NOT "
                 "guaranteed at this size - a 256 MB buffer is written between timed calls"}


def run_abc(ctx, args, N, with_e2e, with_cpu, flush_small=True):
    """One abc() workload at N global rows -> the JSON line (dict)."""
    torch, dist = ctx.torch, ctx.dist
    from abc_dls_b200 import synth
    eng, dev, world, rank = ctx.eng, ctx.dev, ctx.world, ctx.rank
    r0, n_local = ctx.shard(N)
    d, P = args.stats, args.params
    par, ss, tgt = synth.abc_tables(n_local, d, P, dev, row_start=r0)
    if args.param_layout == "row":
        par = par.t().contiguous().t()          # (P, n) view of an (n, P) C-order table, strides (1, P)
    small = N * d * 8 <= 3e8                    # the table fits L2: evict it between timed calls
    flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev) if small and flush_small else None

    def step():
        if flush is not None:
            flush.add_(1)
        return eng.abc(tgt, par, ss, args.tol, args.method, shard=(N, r0))

    ms_step, launches, clocks, res = timed_region(ctx, args, step)
    if os.environ.get("ABCDLS_PEER_TRACE") == "1" and world > 1:   # timeline of the exchanges of the timed calls
        tr = eng.comm.trace()
        if tr is not None:
            import numpy as np
            os.makedirs("gpurun_out", exist_ok=True)
            np.save(f"gpurun_out/peer_trace_r{rank}.npy", tr)
    if flush is not None:   # the flush kernel itself is inside the timed region: measure and subtract it
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.sync()
        ev0.record()
        for _ in range(args.steps):
            flush.add_(1)
        ev1.record()
        ctx.sync()
        ms_step = max(ms_step - ctx.allmax(ev0.elapsed_time(ev1)) / args.steps, 1e-6)
    value = N / (ms_step * 1e-3)
    digest = ctx.digest(res.idx, res.unadj_values)
    vals = res.adj_values if res.adj_values is not None else res.unadj_values
    digest["values_sum"] = ctx.allsum(float(vals.sum().item()))
    digest["n_accepted"] = int(ctx.allsum(float(res.idx.numel())))
    kern = stage_times(ctx, args, step)
    roofline = streaming_roofline(kern, d, n_local, 16 * d + 16, N, ms_step, world)

    # ---- end to end through host buffers -----------------------------------------------------------
    e2e = None
    if with_e2e:
        if args.param_layout == "row":
            h_par = torch.empty((n_local, P), dtype=torch.float64, pin_memory=True).t()
        else:
            h_par = torch.empty((P, n_local), dtype=torch.float64, pin_memory=True)
        h_ss = torch.empty((d, n_local), dtype=torch.float64, pin_memory=True)
        h_par.copy_(par)
        h_ss.copy_(ss)
        h_tgt = tgt.cpu().pin_memory()
        h_out = torch.empty((P, res.nacc), dtype=torch.float64, pin_memory=True)   # posterior sample, (P, n_acc)
        h_stats = torch.empty((P, 6), dtype=torch.float64, pin_memory=True)
        d2h, h2d_dyn = [0], [0]
        parts = []

        def e2e_step(src=None):
            # sumstat must be resident; param stays in pinned host memory and only the accepted rows are read by the
            # gather kernel through UVA (counted below as H2D traffic)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            src = h_ss if src is None else src
            dss = src.to(dev, non_blocking=True)
            if dss.dtype != torch.float64:
                dss = dss.to(torch.float64)        # float32 host table: half the PCIe bytes, widened (exactly) in HBM
            dt_ = h_tgt.to(dev, non_blocking=True)
            ev[1].record()
            r = eng.abc(dt_, h_par, dss, args.tol, args.method, shard=(N, r0))
            ev[2].record()
            v_ = (r.adj_values if r.adj_values is not None else r.unadj_values).t()   # (P, n_local_acc), contiguous
            n_loc = v_.shape[1]
            # pinned AND dense destination: PCIe line rate (a strided pinned view is copied at ~2 GB/s)
            h_out.view(-1)[: P * n_loc].view(P, n_loc).copy_(v_, non_blocking=True)
            h_stats.copy_(r.stats, non_blocking=True)
            ev[3].record()
            torch.cuda.current_stream(dev).synchronize()
            parts.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
            d2h[0] = n_loc * P * 8 + h_stats.numel() * 8
            h2d_dyn[0] = (src.numel() * src.element_size() + h_tgt.numel() * 8       # gathered parameters: one 8 P-byte
                          + r.idx.numel() * P * (8 if args.param_layout == "row" else 32))  # record / 32-byte sectors
            return h_out, h_stats

        del par, ss
        torch.cuda.empty_cache()
        for _ in range(3):     # like the device-resident leg: call 1 eager, call 2 records the CUDA graph, call 3 replays
            e2e_step()
        ctx.sync()
        nstep = max(1, min(args.steps, 3))
        del parts[:]
        t0 = time.perf_counter()
        for _ in range(nstep):
            e2e_step()
        ctx.sync()
        dt_e = ctx.allmax((time.perf_counter() - t0) / nstep)
        e2e = {"value": N / dt_e, "unit": UNIT, "h2d_bytes_per_step": h2d_dyn[0], "d2h_bytes_per_step": d2h[0],
               "ms_per_step": dt_e * 1e3, "steps": nstep,
               "breakdown_ms": dict(zip(("h2d_sumstat", "abc_incl_zero_copy_param_gather", "d2h_results"),
                                        [round(sum(p_[i] for p_ in parts) / len(parts), 3) for i in range(3)])),
               "api": "AbcEngine.abc(target, param, sumstat) on pinned HOST tables: sumstat H2D (PCIe line rate), param "
                      "gathered zero-copy (accepted rows only) -> posterior sample + summary D2H into pinned buffers",
               "host_numa_binding": ctx.numa}
        if not args.no_e2e_f32:
            # The same call when the caller hands over float32 (what ABC-DLS does: Keras predictions, ABC.py:1849; the
            # synthetic values are float32-representable by construction): identical results, half the H2D.
            h2d_main = h2d_dyn[0]
            h_ss32 = torch.empty((d, n_local), dtype=torch.float32, pin_memory=True)
            h_ss32.copy_(h_ss)
            exact = bool((h_ss32[:, :1 << 20].double() == h_ss[:, :1 << 20]).all())
            for _ in range(3):
                e2e_step(h_ss32)
            ctx.sync()
            del parts[:]
            t0 = time.perf_counter()
            for _ in range(nstep):
                e2e_step(h_ss32)
            ctx.sync()
            dt32 = ctx.allmax((time.perf_counter() - t0) / nstep)
            e2e["float32_host_tables"] = {
                "value": N / dt32, "ms_per_step": dt32 * 1e3, "h2d_bytes_per_step": h2d_dyn[0],
                "values_identical_to_f64_tables": exact,
                "breakdown_ms": dict(zip(("h2d_sumstat_and_widen", "abc_incl_zero_copy_param_gather", "d2h_results"),
                                         [round(sum(p_[i] for p_ in parts) / len(parts), 3) for i in range(3)]))}
            h2d_dyn[0] = h2d_main
            del h_ss32
        del h_par, h_ss, h_out

    base = None
    if with_cpu and rank == 0:
        base, _ = cpu_arm(args, 2, 1, "abc")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": abc_config(args, world, N),
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "nacc": res.nacc, "ds": res.ds, "digest": digest}
    line.update(engine_facts(ctx))
    return line


def run_postpr(ctx, args):
    """BASELINE cfg2: postpr(rejection) over N x K softmax summaries rounded to 5 decimals (real distance ties).
    SURVEY 8d bytes per simulation: 2 x 8 K + 16 + 4 (model id) = 68 at K = 3."""
    torch = ctx.torch
    from abc_dls_b200 import synth
    eng, dev, world, rank = ctx.eng, ctx.dev, ctx.world, ctx.rank
    N, K = int(args.rows), args.stats
    r0, n_local = ctx.shard(N)
    ids, ss, tgt = synth.postpr_tables(n_local, dev, K=K, row_start=r0)
    models = [f"M{k_}" for k_ in range(K)]
    flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)     # 72 MB table: evict it from L2 between calls

    def step():
        flush.add_(1)
        return eng.postpr(tgt, ids, models, ss, args.tol, "rejection", shard=(N, r0))

    ms_step, launches, clocks, res = timed_region(ctx, args, step)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    ev0.record()
    for _ in range(args.steps):
        flush.add_(1)
    ev1.record()
    ctx.sync()
    ms_step = max(ms_step - ctx.allmax(ev0.elapsed_time(ev1)) / args.steps, 1e-6)
    kern = stage_times(ctx, args, step)
    bps = 16 * K + 16 + 4
    roofline = streaming_roofline(kern, K, n_local, bps, N, ms_step, world)
    digest = ctx.digest(res.idx, res.values.to(torch.float64))
    digest["counts"] = [int(c) for c in res.counts.cpu().tolist()]

    # end to end: float64 summaries + int32 model ids from pinned host memory, the K counts back
    e2e = None
    if not args.no_e2e:
        h_ss, h_ids, h_tgt = ss.cpu().pin_memory(), ids.cpu().pin_memory(), tgt.cpu().pin_memory()
        h_cnt = torch.empty(K, dtype=torch.int64, pin_memory=True)

        def e2e_step():
            dss, did = h_ss.to(dev, non_blocking=True), h_ids.to(dev, non_blocking=True)
            r = eng.postpr(h_tgt.to(dev, non_blocking=True), did, models, dss, args.tol, "rejection", shard=(N, r0))
            h_cnt.copy_(r.counts, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return h_cnt
        for _ in range(3):
            e2e_step()
        ctx.sync()
        nstep = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(nstep):
            e2e_step()
        ctx.sync()
        dt_e = ctx.allmax((time.perf_counter() - t0) / nstep)
        e2e = {"value": N / dt_e, "unit": UNIT, "ms_per_step": dt_e * 1e3, "steps": nstep,
               "h2d_bytes_per_step": h_ss.numel() * 8 + h_ids.numel() * 4 + K * 8, "d2h_bytes_per_step": K * 8,
               "api": "AbcEngine.postpr(target, model ids, sumstat) on pinned HOST tables -> H2D -> counts D2H"}
    base = None
    if not args.no_cpu and rank == 0:
        base, _ = cpu_arm(args, 2, 1, "postpr")
    line = {"metric": "simulations scored/sec (postpr rejection)", "value": N / (ms_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg2 postpr(method=rejection): N={N} sims x K={K} models (softmax summaries rounded "
                                   f"to 5 decimals: real ties), tol={args.tol}", "rows": N, "models": K, "tol": args.tol,
                       "sharding": f"rows/{world}", "l2": "256 MB buffer written between timed calls (the 72 MB table "
                                                         "would otherwise stay in L2); its time is subtracted"},
            "roofline": roofline, "cpu_baseline": base, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "nacc": res.nacc, "ds": res.ds, "digest": digest}
    line.update(engine_facts(ctx))
    return line


def run_smc(ctx, args):
    """BASELINE cfg4: `rounds` SMC rounds.  Per round (timed): per-column abc(rejection) for all P parameters
    (ABC.py:2805-2840) -> posterior min / max -> oldrange over the simulated parameters (2714-2716) -> updating_newrange
    + noise_injection_update (decrease .95, increase .01, hard range) -> box filter of the simulated rows (2974-2990).
    Between rounds (untimed, the simulator's job): N new simulations inside the new box, seed SEED + 1000 round."""
    torch = ctx.torch
    import numpy as np
    from abc_dls_b200 import synth, smc
    eng, dev, world, rank = ctx.eng, ctx.dev, ctx.world, ctx.rank
    N, P = int(args.rows), args.params
    r0, n_local = ctx.shard(N)
    hlo, hhi = synth.param_ranges(P, dev)
    hmin, hmax = hlo.cpu().numpy(), hhi.cpu().numpy()
    _, _, tgt = synth.abc_tables(16, P, P, dev)          # the observation stays put while the box shrinks
    flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)

    def job(timers=None):
        lo, hi = hlo.clone(), hhi.clone()
        t_hot, hist = 0.0, []
        for rnd in range(args.rounds):
            par, ss, _ = synth.abc_tables(n_local, P, P, dev, row_start=r0, seed=synth.SEED + 1000 * rnd, lo=lo, hi=hi)
            flush.add_(1)
            ctx.sync()
            t0 = time.perf_counter()
            s = eng.abc_columns_summary(tgt, par, ss, args.tol, "rejection", shard=(N, r0), want=("min", "max"))
            omin, omax = eng.oldrange(par)
            omin_h, omax_h = omin.cpu().numpy(), omax.cpu().numpy()
            rng = smc.updating_newrange(s["min"], s["max"], omin_h, omax_h, 0.95)
            rng = smc.noise_injection_update(rng, omin_h, omax_h, 0.01, hmin, hmax, 0.95)
            keep = eng.box_filter(par, torch.from_numpy(rng.min), torch.from_numpy(rng.max), shard=(N, r0))
            torch.cuda.synchronize(dev)
            t_hot += time.perf_counter() - t0
            hist.append({"round": rnd, "lmrd": smc.lmrd_calculation(rng, hmin, hmax),
                         "kept_rows": int(ctx.allsum(float(keep.numel()))),
                         "rows_sum_u64": ctx.digest(keep, keep.to(torch.float64))["rows_sum_u64"]})
            lo, hi = torch.from_numpy(rng.min).to(dev), torch.from_numpy(rng.max).to(dev)
            del par, ss, keep
        return ctx.allmax(t_hot), hist

    for _ in range(max(1, min(args.warmup, 2))):
        job()
    l0 = eng.k.launches
    sampler = ClockSampler(ctx.local_rank, args.clock_interval)
    nstep = max(1, min(args.steps, 5))
    with sampler as clocks:
        runs = [job() for _ in range(nstep)]
    launches = (eng.k.launches - l0) // nstep
    t_job = sum(r[0] for r in runs) / nstep
    hist = runs[-1][1]
    sims = N * args.rounds
    # stage times of one round at full size
    par, ss, _ = synth.abc_tables(n_local, P, P, dev, row_start=r0)
    kern = stage_times(ctx, args, lambda: eng.abc_columns_summary(tgt, par, ss, args.tol, "rejection", shard=(N, r0),
                                                                  want=("min", "max")))
    bps = 16 * P + 8 * P + 8 * P      # per-column mode: statistics table once + oldrange + box filter over the parameters
    roofline = streaming_roofline(kern, P, n_local, bps, sims, t_job * 1e3, world)
    base = None
    if not args.no_cpu and rank == 0:
        base, _ = cpu_arm(args, 1, 0, "smc")
    line = {"metric": "simulations scored/sec (SMC rounds: per-column ABC rejection + range update + box filter)",
            "value": sims / t_job, "unit": UNIT, "n_gpus": world, "steps": nstep, "warmup": max(1, min(args.warmup, 2)),
            "ms_per_step": t_job * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"cfg4 Run_SMC: {args.rounds} rounds x N={N} sims/round x P={P} parameters, per-column "
                                   f"abc(rejection, tol={args.tol}) -> min/max -> updating_newrange + "
                                   f"noise_injection_update (decrease .95, increase .01, hardrange) -> box filter; a step "
                                   f"is the whole job, the regeneration of the tables between rounds is not timed",
                       "rows": N, "params": P, "rounds": args.rounds, "tol": args.tol, "sharding": f"rows/{world}",
                       "l2": "tables regenerated and a 256 MB buffer written between rounds"},
            "timing": "host wall clock around each round's hot path (device synchronised on both sides), summed over "
                      "the rounds, max over ranks: the range rules are host code between device passes",
            "roofline": roofline, "cpu_baseline": base, "e2e": None, "gpu_launches": launches,
            "clocks": clocks.summary(), "rounds": hist, "ms_per_round": t_job * 1e3 / args.rounds}
    line.update(engine_facts(ctx))
    return line


def main():
    args = _args()
    if args.workload == "cv4abc":
        return cv_workload(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        kind = {"abc": "abc", "sweep": "abc", "postpr": "postpr", "smc": "smc"}[args.workload]
        base, dt = cpu_arm(args, max(1, args.steps), max(0, min(args.warmup, 1)), kind)
        cfg = abc_config(args, world) if kind == "abc" else {"workload": f"{args.workload} (see the b200 arm)",
                                                             "rows": int(args.rows), "tol": args.tol}
        cfg["rows_timed"] = base["rows_timed"]
        metric = {"abc": METRIC, "postpr": "simulations scored/sec (postpr rejection)",
                  "smc": "simulations scored/sec (SMC rounds: per-column ABC rejection + range update + box filter)"}[kind]
        line = {"metric": metric, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "impl": "reference",
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    ctx = Ctx()
    if args.workload == "abc":
        line = run_abc(ctx, args, int(args.rows), not args.no_e2e, not args.no_cpu)
    elif args.workload == "postpr":
        line = run_postpr(ctx, args)
    elif args.workload == "smc":
        line = run_smc(ctx, args)
    else:   # sweep: cfg5 at every size of BASELINE config 5; the headline line is the largest size
        sizes = [int(1e6), int(3e6), int(1e7), int(3e7), int(1e8)]
        entries = []
        for n in sizes:
            ln = run_abc(ctx, args, n, False, False)
            entries.append({k_: ln[k_] for k_ in ("value", "ms_per_step", "nacc", "digest", "front_half")}
                           | {"rows": n, "whole_call_frac_of_peak_per_gpu": ln["roofline"]["whole_call"]["frac_of_peak_per_gpu"],
                              "kernel": ln["roofline"]["kernel"], "kernel_frac": ln["roofline"]["frac"]})
            ctx.torch.cuda.empty_cache()
        line = ln
        line["sweep"] = entries
    if rank == 0:
        print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
