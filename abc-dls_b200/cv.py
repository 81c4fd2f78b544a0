"""Leave-one-out tools of R's abc package on the CUDA engine: ``cv4abc``, ``cv4postpr``, ``gfit``.

Reference call sites: src/Classes/ABC.py:1900-1902, 1913 (``abc.cv4abc(param=, sumstat=, nval=, tols=, method=, trace=)``
then ``summary(...)``), ABC.py:886-892 (``abc.cv4postpr(index=, sumstat=, nval=, tol=, method=)``), ABC.py:961-963
(``abc.gfit(target, ss, repeats, tol=)``).  They run before the final estimate in every ``After_train`` and cost
``nval`` (100) abc()/postpr() calls each, per parameter column.

R removes the held-out row physically (``param[-i, ]``, ``sumstat[-i, ]``): N - 1 rows enter the MADs, nacc =
ceiling((N - 1) tol), and ties keep R's row order.  Here the N - 1 row table lives in ONE device buffer whose address
never changes; moving the hole from row a to row s rewrites only the rows between them (``|s - a|`` rows per column,
one strided copy), so a whole cross-validation moves about one table's worth of bytes when the rows are visited in
ascending order, and every validation after the second replays the engine's CUDA graph (same key: same buffers, same
shapes).  Point estimates stay on the device until the end.

Per-column mode (one parameter, one statistic - how ABC-DLS calls cv4abc for rejection / loclinear, ABC.py:1896-1904)
on a table that fits one CTA's shared memory does not loop at all: every (tolerance, held-out row) fit is one thread
block of ``csrc/small.cu`` and the whole cross-validation is ONE launch per tolerance.

Single rank only: the tables of these steps are test-set sized in ABC-DLS (``test_size`` rows).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .engine import AbcEngine, AbcError
from . import summary as _summary


class LeaveOneOut:
    """Device copies of (cols, N) tables with one row left out, kept in place between validations."""

    def __init__(self, tables: Sequence[torch.Tensor]):
        self.src = list(tables)
        self.n = self.src[0].shape[1]
        if self.n < 3:
            raise AbcError("leave-one-out needs at least 3 rows")
        self.hole = 0
        self.buf = []
        pitch = self.n + (self.n & 1)                  # even pitch: keeps 16-byte column alignment for the TMA path
        for t in self.src:
            b = torch.empty((t.shape[0], pitch), dtype=t.dtype, device=t.device)[:, : self.n - 1]
            b.copy_(t[:, 1:])
            self.buf.append(b)

    def hold_out(self, s: int) -> List[torch.Tensor]:
        """Buffers = the tables without row ``s`` (row order kept)."""
        a = self.hole
        if not (0 <= s < self.n):
            raise AbcError("held-out row out of range")
        for t, b in zip(self.src, self.buf):
            if s > a:
                b[:, a:s].copy_(t[:, a:s])             # rows [a, s) move back to their own position
            elif s < a:
                b[:, s:a].copy_(t[:, s + 1:a + 1])     # rows (s, a] shift up by one
        self.hole = s
        return self.buf


def _draw(n: int, k: int, seed) -> np.ndarray:
    if k > n:
        raise AbcError("cannot take a sample larger than the population")      # R's sample() error
    return np.random.default_rng(seed).choice(n, size=k, replace=False).astype(np.int64)


def _engine_single(eng: AbcEngine) -> None:
    if eng.comm.world > 1:
        raise AbcError("cv4abc / cv4postpr / gfit run on one rank (their tables are test-set sized)")


class Cv4abcResult:
    """Fields of R's class ``cv4abc``: cvsamples (0-based rows), tols (ascending), true (nval x P), estim[tol]."""

    def __init__(self, cvsamples, tols, true, estim, names, statistic):
        self.cvsamples, self.tols, self.true, self.estim = cvsamples, tols, true, estim
        self.names, self.statistic = names, statistic

    def prediction_error(self) -> np.ndarray:
        """summary.cv4abc: colSums((estim - true)^2) / (var(true) * nval), one row per tolerance."""
        nval = self.true.shape[0]
        truevar = self.true.var(axis=0, ddof=1) * nval
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.stack([((self.estim[float(t)] - self.true) ** 2).sum(axis=0) / truevar for t in self.tols])

    def summary(self, print_: bool = True) -> np.ndarray:
        tab = self.prediction_error()
        if print_:
            print(self.summary_text())
        return tab

    def summary_text(self, digits: int = 4) -> str:
        tab = self.prediction_error()
        names = self.names["parameter.names"]
        cells = [[f"{v:.{digits}f}" for v in row] for row in tab]
        labels = [f"{t:g}" for t in self.tols]
        lab = max(len(x) for x in labels)
        widths = [max(len(names[p]), *(len(c[p]) for c in cells)) for p in range(len(names))]
        lines = [f"Prediction error based on a cross-validation sample of {self.true.shape[0]}", "",
                 " " * lab + " " + " ".join(n.rjust(w) for n, w in zip(names, widths))]
        for l, c in zip(labels, cells):
            lines.append(l.ljust(lab) + " " + " ".join(x.rjust(w) for x, w in zip(c, widths)))
        return "\n".join(lines)


def _small_status(st: int, method: str) -> None:
    """The R stop() conditions of abc() for a one-CTA fit (same messages as engine.abc)."""
    from . import _capi
    if st & _capi.S_NONFINITE:
        raise AbcError("non-finite values in 'sumstat'/'target' (R's NA-row handling is not implemented)")
    if st & _capi.S_ZERO_VAR:
        raise AbcError("Zero variance in the summary statistics.")
    if st & _capi.S_ZERO_VAR_REGION:
        msg = "Zero variance in the summary statistics in the selected region."
        if method != "rejection":
            raise AbcError(msg + " Try: checking summary statistics, choosing larger tolerance, or rejection method.")
        print("Warning message:\n" + msg + " Check summary statistics, consider larger tolerance.")
    if st & _capi.S_SINGULAR:
        raise AbcError("lsfit: X matrix deemed to be singular (collinear summary statistics in the accepted region)")


def small_fits(eng: AbcEngine, par: torch.Tensor, ss: torch.Tensor, col, hold, tol: float, method: str,
               hcorr: bool = True, target=None) -> Optional[np.ndarray]:
    """B one-dimensional abc() problems in one launch (csrc/small.cu): problem b uses column col[b] of ``par`` / ``ss``
    with row hold[b] left out and the held-out statistic as target (hold None: no row removed, target[b] given).
    Returns (B, 8) = {median, mean, ds, mad, min, max, n_acc, status} or None when the table does not qualify."""
    k = eng.k
    N = ss.shape[1]
    n1 = N - (0 if hold is None else 1)
    nacc = int(math.ceil(n1 * tol))
    if not (1 <= nacc <= n1):
        raise AbcError("'tol' must give 1 <= ceiling(N * tol) <= N")
    if not (hasattr(k, "small_columns") and ss.stride(1) == 1 and k.small_supported(N, nacc)):
        return None
    dev = eng.device
    col_t = torch.as_tensor(np.asarray(col, dtype=np.int32)).to(dev)
    hold_t = None if hold is None else torch.as_tensor(np.asarray(hold, dtype=np.int64)).to(dev)
    tgt_t = None if target is None else target.to(device=dev, dtype=torch.float64).contiguous()
    out = k.new(col_t.numel(), 8)
    k.small_columns(ss, par, N, col_t, hold_t, tgt_t, nacc, method == "loclinear", hcorr, out)
    o = out.cpu().numpy()
    st = 0
    for v in np.unique(o[:, 7]):
        st |= int(v)
    _small_status(st, method)
    return o


def cv4abc(eng: AbcEngine, par: torch.Tensor, ss: torch.Tensor, nval: int, tols, method: str,
           statistic: str = "median", hcorr: bool = True, cvsamples=None, seed=None, names=None) -> Cv4abcResult:
    """par (P, N), ss (d, N): float64 device tables.  One abc() per (tolerance, held-out row)."""
    _engine_single(eng)
    if statistic not in ("median", "mean", "mode"):
        raise AbcError("Statistic has to be mean, median or mode.")
    P, N = par.shape
    samp = _draw(N, int(nval), seed) if cvsamples is None else np.asarray(cvsamples, dtype=np.int64).reshape(-1)
    tols = np.sort(np.atleast_1d(np.asarray(tols, dtype=np.float64)))
    names = names or {"parameter.names": [f"P{p + 1}" for p in range(P)]}
    true = par[:, torch.from_numpy(samp).to(par.device)].t().cpu().numpy()
    estim: Dict[float, np.ndarray] = {}
    if P == 1 and ss.shape[0] == 1 and statistic in ("median", "mean") and method in ("rejection", "loclinear"):
        # per-column mode on a small table: every fit is one thread block, one launch per tolerance
        for tol in tols:
            o = small_fits(eng, par, ss, np.zeros(samp.size, dtype=np.int32), samp, float(tol), method, hcorr)
            if o is None:
                estim.clear()
                break
            estim[float(tol)] = o[:, 0 if statistic == "median" else 1].reshape(-1, 1).copy()
        if estim:
            return Cv4abcResult(samp, tols, true, estim, names, statistic)
    loo = LeaveOneOut([par, ss])
    order = np.argsort(samp, kind="stable")            # ascending holes: every row of the buffers moves at most once
    for tol in tols:
        est = torch.empty((samp.size, P), dtype=torch.float64, device=eng.device)
        for i in order:
            s = int(samp[i])
            bpar, bss = loo.hold_out(s)
            res = eng.abc(ss[:, s], bpar, bss, float(tol), method, hcorr=hcorr)
            est[i] = _summary.point_estimate(res.values, res.weights, res.stats, res.adj_values is not None, statistic)
        estim[float(tol)] = est.cpu().numpy()
    return Cv4abcResult(samp, tols, true, estim, names, statistic)


def cv4abc_columns(eng: AbcEngine, par: torch.Tensor, ss: torch.Tensor, nval: int, tol: float, method: str,
                   statistic: str = "median", hcorr: bool = True, cvsamples=None, seed=None, names=None):
    """The whole "Cv error independently" loop of ABC-DLS (ABC.py:1896-1904: one cv4abc per parameter column, column c of
    ``ss`` against column c of ``par``) as ONE launch of P x nval thread blocks when the table is small; otherwise a loop
    over cv4abc.  Returns one Cv4abcResult per column (the same held-out rows for every column)."""
    _engine_single(eng)
    P, N = par.shape
    samp = _draw(N, int(nval), seed) if cvsamples is None else np.asarray(cvsamples, dtype=np.int64).reshape(-1)
    pn = (names or {}).get("parameter.names", [f"P{p + 1}" for p in range(P)])
    o = None
    if statistic in ("median", "mean") and method in ("rejection", "loclinear") and ss.shape[0] == P:
        o = small_fits(eng, par, ss, np.repeat(np.arange(P, dtype=np.int32), samp.size), np.tile(samp, P), float(tol),
                       method, hcorr)
    if o is None:
        return [cv4abc(eng, par[c:c + 1], ss[c:c + 1], nval, tol, method, statistic, hcorr, samp, None,
                       {"parameter.names": [pn[c]]}) for c in range(P)]
    est = o[:, 0 if statistic == "median" else 1].reshape(P, samp.size)
    true = par[:, torch.from_numpy(samp).to(par.device)].cpu().numpy()
    tols = np.array([float(tol)])
    return [Cv4abcResult(samp, tols, true[c].reshape(-1, 1), {float(tol): est[c].reshape(-1, 1).copy()},
                         {"parameter.names": [pn[c]]}, statistic) for c in range(P)]


class Cv4postprResult:
    """Fields of R's class ``cv4postpr``: cvsamples, tols, true (labels), estim[tol] ((nval K) x K probabilities)."""

    def __init__(self, cvsamples, tols, true_ids, models, estim, nval, method="rejection"):
        self.cvsamples, self.tols, self.true_ids, self.models, self.estim = cvsamples, tols, true_ids, models, estim
        self.true = [models[i] for i in true_ids]
        self.nval, self.method = nval, method

    def _tables(self):
        K = len(self.models)
        conf, probs = {}, {}
        for t in self.tols:
            e = self.estim[float(t)]
            cm = np.zeros((K, K), dtype=np.int64)
            np.add.at(cm, (self.true_ids, e.argmax(axis=1)), 1)        # which.max: first model on ties
            conf[float(t)] = cm
            with np.errstate(invalid="ignore"):
                probs[float(t)] = np.stack([e[self.true_ids == k].mean(axis=0) if (self.true_ids == k).any()
                                            else np.full(K, np.nan) for k in range(K)])
        return conf, probs

    def summary(self, print_: bool = True):
        conf, probs = self._tables()
        if print_:
            print(self.summary_text())
        return {"conf.matrix": conf, "probs": probs}

    def summary_text(self, digits: int = 4) -> str:
        conf, probs = self._tables()
        w = max(max(len(m) for m in self.models), digits + 2)
        head = " " * w + " " + " ".join(m.rjust(w) for m in self.models)
        lines = [f"Confusion matrix based on {self.nval} samples for each model.", ""]
        for t in self.tols:
            lines += [f"$tol{t:g}", head]
            lines += [m.ljust(w) + " " + " ".join(str(int(v)).rjust(w) for v in conf[float(t)][i])
                      for i, m in enumerate(self.models)]
            lines.append("")
        lines += ["", f"Mean model posterior probabilities ({self.method})", ""]
        for t in self.tols:
            lines += [f"$tol{t:g}", head]
            lines += [m.ljust(w) + " " + " ".join(f"{v:.{digits}f}".rjust(w) for v in probs[float(t)][i])
                      for i, m in enumerate(self.models)]
            lines.append("")
        return "\n".join(lines)


def cv4postpr(eng: AbcEngine, model_id: torch.Tensor, models: List[str], ss: torch.Tensor, nval: int, tols,
              method: str = "rejection", cvsamples=None, seed=None) -> Cv4postprResult:
    """model_id (N,) int32 device ids into ``models`` (sorted levels), ss (d, N).  ``nval`` held-out rows PER model."""
    _engine_single(eng)
    if method != "rejection":
        raise AbcError("cv4postpr: only method='rejection' is implemented (mnlogistic / neuralnet stay with R's nnet)")
    K, N = len(models), ss.shape[1]
    ids_host = model_id.cpu().numpy().astype(np.int64)
    if cvsamples is None:
        rng = np.random.default_rng(seed)
        parts = []
        for k in range(K):
            rows = np.flatnonzero(ids_host == k)
            if rows.size < nval:
                raise AbcError("cannot take a sample larger than the population")
            parts.append(rng.choice(rows, size=int(nval), replace=False))
        samp = np.concatenate(parts).astype(np.int64)
    else:
        samp = np.asarray(cvsamples, dtype=np.int64).reshape(-1)
    tols = np.sort(np.atleast_1d(np.asarray(tols, dtype=np.float64)))
    ids = model_id.to(device=eng.device, dtype=torch.int32).reshape(1, -1)
    loo = LeaveOneOut([ids, ss])
    order = np.argsort(samp, kind="stable")
    estim: Dict[float, np.ndarray] = {}
    for tol in tols:
        est = torch.empty((samp.size, K), dtype=torch.float64, device=eng.device)
        for i in order:
            s = int(samp[i])
            bid, bss = loo.hold_out(s)
            est[i] = eng.postpr(ss[:, s], bid[0], models, bss, float(tol), method).pred
        estim[float(tol)] = est.cpu().numpy()
    return Cv4postprResult(samp, tols, ids_host[samp], list(models), estim, int(nval), method)


class GfitResult:
    """R's class ``gfit``: dist.obs and the null sample dist.sim."""

    def __init__(self, dist_obs: float, dist_sim: np.ndarray):
        self.dist_obs, self.dist_sim = dist_obs, dist_sim

    def summary(self, print_: bool = True):
        x = np.sort(self.dist_sim)
        q = lambda p: float(_summary._quantile7(torch.from_numpy(x), p))
        out = {"pvalue": float((self.dist_sim > self.dist_obs).sum() / self.dist_sim.size),
               "s.dist.sim": np.array([x[0], q(0.25), q(0.5), float(self.dist_sim.mean()), q(0.75), x[-1]]),
               "dist.obs": self.dist_obs}
        if print_:
            print(self.summary_text(out))
        return out

    def summary_text(self, out: Optional[dict] = None, digits: int = 4) -> str:
        out = out or self.summary(print_=False)
        labs = ["Min.", "1st Qu.", "Median", "Mean", "3rd Qu.", "Max."]
        cells = [f"{v:.{digits}g}" for v in out["s.dist.sim"]]
        w = [max(len(a), len(b)) for a, b in zip(labs, cells)]
        return "\n".join(["$pvalue", f"[1] {out['pvalue']:.{digits}g}", "", "$s.dist.sim",
                          " ".join(a.rjust(x) for a, x in zip(labs, w)), " ".join(c.rjust(x) for c, x in zip(cells, w)),
                          "", "$dist.obs", f"[1] {out['dist.obs']:.{digits}g}"])


def gfit(eng: AbcEngine, target: torch.Tensor, ss: torch.Tensor, nb_replicate: int, tol: float,
         statistic: str = "mean", samples=None, seed=None) -> GfitResult:
    """Goodness of fit: ``statistic`` of the accepted distances (abc rejection) for the observation, and its null
    distribution from ``nb_replicate`` simulations each held out and used as the observation."""
    _engine_single(eng)
    if statistic not in ("mean", "median"):
        raise AbcError("gfit: statistic must be mean or median")
    d, N = ss.shape

    def stat(res) -> torch.Tensor:
        x = res.dist_accepted
        if statistic == "mean":
            return x.mean()
        return _summary._quantile7(torch.sort(x).values, 0.5)

    samp = _draw(N, int(nb_replicate), seed) if samples is None else np.asarray(samples, dtype=np.int64).reshape(-1)
    obs = float(stat(eng.abc(target, ss[:1], ss, float(tol), "rejection")))
    loo = LeaveOneOut([ss])
    sim = torch.empty(samp.size, dtype=torch.float64, device=eng.device)
    for i in np.argsort(samp, kind="stable"):
        s = int(samp[i])
        (bss,) = loo.hold_out(s)
        sim[i] = stat(eng.abc(ss[:, s], bss[:1], bss, float(tol), "rejection"))
    return GfitResult(obs, sim.cpu().numpy())
