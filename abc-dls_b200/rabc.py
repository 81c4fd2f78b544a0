"""Drop-in for the rpy2 handle ABC-DLS binds as the module global ``abc`` (src/Classes/ABC.py:36).

Same attribute names and keyword arguments as the reference's call sites:

    abc.abc(target=, param=, sumstat=, method=, tol=)            ABC.py:1950-1953, 1964, 2806-2809, 2817
    abc.postpr(target=, index=, sumstat=, tol=, method=)         ABC.py:919
    abc.cv4abc(param=, sumstat=, nval=, tols=, method=, trace=)  ABC.py:1900-1902, 1913
    abc.cv4postpr(index=, sumstat=, nval=, tol=, method=)        ABC.py:886
    abc.gfit(target, ss, repeats, tol=)                          ABC.py:961

Inputs may be pandas DataFrame/Series, numpy arrays or torch tensors (host or CUDA).  Host inputs are
copied to HBM once per call; the arithmetic runs in the CUDA engine (``engine.py``) — there is no CPU
path.  R's defaults that matter are kept: ``hcorr=True``, ``kernel='epanechnikov'``, ``transf='none'``.
Errors R raises with ``stop()`` surface as :class:`AbcError` with R's message.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .engine import AbcEngine, AbcError, AbcResult, PostprResult
from . import summary as _summary
from . import cv as _cv

_engine: Optional[AbcEngine] = None
# R draws the held-out rows of cv4abc / cv4postpr / gfit from its session RNG (set.seed); the reference never seeds it.
# ``default_seed`` plays set.seed's role for callers that cannot pass ``seed=`` (ABC-DLS's unmodified call sites).
default_seed: Optional[int] = None


def engine() -> AbcEngine:
    """Process-wide engine on the current CUDA device (created on first use)."""
    global _engine
    if _engine is None:
        _engine = AbcEngine()
    return _engine


def set_engine(e: Optional[AbcEngine]) -> None:
    global _engine
    _engine = e


def _names(x, n: int, prefix: str) -> List[str]:
    cols = getattr(x, "columns", None)
    if cols is not None:
        return [str(c) for c in cols]
    name = getattr(x, "name", None)
    if name is not None and n == 1:
        return [str(name)]
    return [f"{prefix}{i + 1}" for i in range(n)]


def to_table(x, device, row_major_ok: bool = False) -> torch.Tensor:
    """(rows, cols) host/device data -> (cols, rows) float64 CUDA table with unit stride along rows.
    A column-major host array (pandas' usual block layout, R's layout) is copied without a transpose.  float32 input
    (what ABC-DLS hands over: Keras predictions, ABC.py:1849, 2758) crosses PCIe as float32 and is widened on the device
    - exact, and half the bytes of rpy2's host-side conversion to R doubles.  ``row_major_ok`` (the parameter table): a
    row-major array is NOT transposed; the (cols, rows) view with strides (1, cols) goes to the engine, whose gather
    reads the accepted rows as contiguous records."""
    if torch.is_tensor(x):
        t = x
        if t.dim() == 1:
            t = t[:, None]
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        t = t.to(device=device, non_blocking=True).to(torch.float64)
        tt = t.t()
        if tt.stride(1) == 1 or (row_major_ok and tt.stride(0) == 1):
            return tt
        return tt.contiguous()
    if hasattr(x, "to_numpy"):
        x = x.to_numpy()
    a = np.asarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    if a.ndim == 1:
        a = a[:, None]
    if a.flags.f_contiguous:
        return torch.from_numpy(a.T).to(device).to(torch.float64)        # already (cols, rows) C-order: one H2D copy
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device).to(torch.float64).t()
    return t if row_major_ok else t.contiguous()                          # + device transpose


def _to_vec(x, device) -> torch.Tensor:
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float64).reshape(-1)
    if hasattr(x, "to_numpy"):
        x = x.to_numpy()
    return torch.from_numpy(np.asarray(x, dtype=np.float64).reshape(-1).copy()).to(device)


class AbcFit:
    """What ``abc.abc`` returns: the fields of R's S3 class ``abc`` as CUDA tensors (accepted rows in
    ascending row order) plus ``summary()``."""

    def __init__(self, res: AbcResult, eng: AbcEngine, parnames: List[str], statnames: List[str]):
        self.res, self._eng = res, eng
        self.names = {"parameter.names": parnames, "statistics.names": statnames}
        self.method, self.numparam, self.numstat = res.method, res.numparam, res.numstat
        self.adj_values, self.unadj_values = res.adj_values, res.unadj_values
        self.ss, self.weights, self.residuals = res.ss, res.weights, res.residuals
        self.dist, self.region_idx = res.dist, res.idx
        self.aic, self.bic = res.aic, res.bic

    def _global_sample(self):
        return self._eng.gather_sample(self.res.values, self.res.weights)

    def summary(self, intvl: float = 0.95, print_: bool = True) -> _summary.SummaryTable:
        vals, w = self._global_sample()
        weighted = self.res.adj_values is not None
        tab = _summary.summary_table(vals, w, self.res.stats, weighted, intvl, self.names["parameter.names"], eng=self._eng)
        if print_:
            print(_summary.summary_text(tab, vals.shape[0], weighted))
        return tab

    def summary_text(self, intvl: float = 0.95) -> str:
        vals, _ = self._global_sample()
        return _summary.summary_text(self.summary(intvl, print_=False), vals.shape[0], self.res.adj_values is not None)


class PostprFit:
    def __init__(self, res: PostprResult):
        self.res = res
        self.models, self.values, self.ss = res.models, res.values, res.ss
        self.pred = res.pred
        self.method = "rejection"

    def _tables(self):
        pred = self.res.pred.cpu().numpy()
        with np.errstate(divide="ignore", invalid="ignore"):
            bf = pred[:, None] / pred[None, :]
        return pred, bf

    def summary(self, print_: bool = True):
        pred, bf = self._tables()
        if print_:
            print(self.summary_text())
        return {"Prob": dict(zip(self.models, pred.tolist())), "BayesF": bf}

    def summary_text(self, digits: int = 4) -> str:
        """summary.postpr as R prints it: the proportions are one named vector (one common format), the Bayes factors a
        matrix (one format per column); digits = max(3, getOption("digits") - 3) = 4."""
        pred, bf = self._tables()
        pcell = _summary.format_r_column(pred, digits)
        w = max(max(len(m) for m in self.models), max(len(c) for c in pcell))
        bcols = [_summary.format_r_column(bf[:, k], digits) for k in range(len(self.models))]
        bw = [max(len(self.models[k]), max(len(c) for c in bcols[k])) for k in range(len(self.models))]
        lab = max(len(m) for m in self.models)
        lines = ["Data:", f" postpr.out$values ({int(self.res.counts.sum())} posterior samples)", "Models a priori:",
                 " " + ", ".join(self.models), "Models a posteriori:", " " + ", ".join(self.models), "",
                 "Proportion of accepted simulations (rejection):",
                 " ".join(m.rjust(w) for m in self.models), " ".join(c.rjust(w) for c in pcell), "",
                 "Bayes factors:", " " * lab + " " + " ".join(m.rjust(x) for m, x in zip(self.models, bw))]
        for i, m in enumerate(self.models):
            lines.append(m.ljust(lab) + " " + " ".join(bcols[k][i].rjust(bw[k]) for k in range(len(self.models))))
        return "\n".join(lines)


_r_abc = None


def _delegate(fn: str, why: str, **kwargs):
    """Arguments this engine does not cover (R's nnet-based methods ``neuralnet`` / ``ridge`` / ``mnlogistic``,
    ``transf``, ``subset``, other kernels) go to the R package the reference binds (ABC.py:36), when rpy2 can load it:
    ``rabc`` then really is a drop-in for the rpy2 handle - the default ``--method`` of two Run_ParamsEstimation
    sub-commands is ``neuralnet``.  Without R the call stops with the reason."""
    global _r_abc
    if _r_abc is None:
        try:
            from rpy2.robjects.packages import importr
            from rpy2.robjects import pandas2ri
            pandas2ri.activate()
            _r_abc = importr("abc")
        except Exception as e:
            raise AbcError(f"{fn}: {why} is not implemented by the B200 engine and R's abc package could not be loaded "
                           f"through rpy2 to take the call ({type(e).__name__}: {e})") from e
    return getattr(_r_abc, fn)(**{k: v for k, v in kwargs.items() if v is not None})


def abc(target, param, sumstat, tol, method, hcorr: bool = True, transf: str = "none", subset=None,
        kernel: str = "epanechnikov", **extra) -> AbcFit:
    plain_transf = transf == "none" or (isinstance(transf, (list, tuple)) and all(t == "none" for t in transf))
    if method not in ("rejection", "loclinear") or not plain_transf or subset is not None or kernel != "epanechnikov":
        why = (f"method='{method}'" if method not in ("rejection", "loclinear") else
               "transf" if not plain_transf else "subset" if subset is not None else f"kernel='{kernel}'")
        return _delegate("abc", why, target=target, param=param, sumstat=sumstat, tol=tol, method=method, hcorr=hcorr,
                         transf=transf, subset=subset, kernel=kernel, **extra)
    eng = engine()
    ss = to_table(sumstat, eng.device)
    par = to_table(param, eng.device, row_major_ok=True)
    tgt = _to_vec(target, eng.device)
    res = eng.abc(tgt, par, ss, float(tol), method, hcorr=hcorr, kernel=kernel)
    for wmsg in res.warnings:
        print("Warning message:\n" + wmsg)
    return AbcFit(res, eng, _names(param, par.shape[0], "P"), _names(sumstat, ss.shape[0], "S"))


def postpr(target, index, sumstat, tol, method="rejection", subset=None, **extra) -> PostprFit:
    if method != "rejection" or subset is not None:
        return _delegate("postpr", f"method='{method}'" if method != "rejection" else "subset", target=target,
                         index=index, sumstat=sumstat, tol=tol, method=method, subset=subset, **extra)
    eng = engine()
    ss = to_table(sumstat, eng.device)
    tgt = _to_vec(target, eng.device)
    ids, models = _model_ids(index, eng)
    res = eng.postpr(tgt, ids, models, ss, float(tol), method)
    return PostprFit(res)


def _model_ids(index, eng: AbcEngine, models=None):
    """R's factor(index): sorted levels + integer codes.  Several ranks: the level table is the union of every rank's
    labels (one object all-gather), so the codes mean the same model everywhere and a level that is absent from one
    shard still has its column in the counts.  ``models`` (explicit level list) skips the discovery."""
    import torch.distributed as dist
    multi = eng.comm.world > 1
    if torch.is_tensor(index) and index.dtype in (torch.int32, torch.int64):
        if models is not None:
            return index, [str(m) for m in models]
        kmax = index.max().to(torch.int64).reshape(1).to(eng.device) if index.numel() else \
            torch.zeros(1, dtype=torch.int64, device=eng.device)
        if multi:
            kmax = kmax.clone()
            eng.comm.allreduce_max(kmax)
        return index, [str(m) for m in range(int(kmax.item()) + 1)]
    labels = np.asarray(index.to_numpy() if hasattr(index, "to_numpy") else index).astype(str)
    if models is None:
        local = np.unique(labels).tolist()
        if multi:
            parts = [None] * eng.comm.world
            dist.all_gather_object(parts, local, group=eng.comm.group)
            local = sorted(set().union(*parts))            # sorted by code point (R sorts in the locale's collation)
        models = local
    models = [str(m) for m in models]
    lut = {m: i for i, m in enumerate(models)}
    try:
        codes = np.fromiter((lut[x] for x in labels), dtype=np.int32, count=labels.size)
    except KeyError as e:
        raise AbcError(f"'index' holds a label that is not among the models: {e}") from None
    return torch.from_numpy(codes), models


def cv4abc(param, sumstat, nval, tols, method, statistic: str = "median", hcorr: bool = True, transf: str = "none",
           subset=None, kernel: str = "epanechnikov", trace: bool = False, abc_out=None, cvsamples=None, seed=None,
           **_ignored) -> _cv.Cv4abcResult:
    """``abc::cv4abc`` (leave-one-out prediction error of the point estimate).  ``cvsamples`` (0-based rows) / ``seed``
    are extras: R draws the rows with its own sample() stream, which is not reproduced."""
    if transf != "none" or subset is not None or kernel != "epanechnikov" or method not in ("rejection", "loclinear"):
        return _delegate("cv4abc", f"method='{method}'" if method not in ("rejection", "loclinear") else
                         "transf / subset / kernel", param=param, sumstat=sumstat, nval=nval, tols=tols, method=method,
                         statistic=statistic, hcorr=hcorr, transf=transf, subset=subset, kernel=kernel, trace=trace)
    eng = engine()
    ss, par = to_table(sumstat, eng.device), to_table(param, eng.device)
    if par.shape[1] != ss.shape[1]:
        raise AbcError("'param' and 'sumstat' must have the same number of rows")
    names = {"parameter.names": _names(param, par.shape[0], "P"), "statistics.names": _names(sumstat, ss.shape[0], "S")}
    return _cv.cv4abc(eng, par, ss, int(nval), tols, method, statistic, hcorr, cvsamples,
                      default_seed if seed is None else seed, names)


def cv4postpr(index, sumstat, nval, tols=None, method="rejection", subset=None, kernel: str = "epanechnikov",
              trace: bool = False, postpr_out=None, cvsamples=None, seed=None, tol=None, **_ignored) -> _cv.Cv4postprResult:
    """``abc::cv4postpr``.  ABC-DLS passes ``tol=`` (ABC.py:886), which R partial-matches to ``tols``."""
    if tols is None:
        tols = tol
    if tols is None:
        raise AbcError("'tols' is missing")
    if subset is not None or method != "rejection":
        return _delegate("cv4postpr", f"method='{method}'" if method != "rejection" else "subset", index=index,
                         sumstat=sumstat, nval=nval, tols=tols, method=method, subset=subset, kernel=kernel, trace=trace)
    eng = engine()
    ss = to_table(sumstat, eng.device)
    ids, models = _model_ids(index, eng)
    if ids.numel() != ss.shape[1]:
        raise AbcError("'index' must be the same length as the number of rows in 'sumstat'.")
    if len(models) < 2:
        raise AbcError("At least two different models must be given.")
    return _cv.cv4postpr(eng, ids, models, ss, int(nval), tols, method, cvsamples, default_seed if seed is None else seed)


def gfit(target, sumstat, nb_replicate, tol, statistic="mean", subset=None, trace: bool = False, samples=None,
         seed=None, **_ignored) -> _cv.GfitResult:
    """``abc::gfit(target, sumstat, nb.replicate, tol, statistic = mean)``."""
    if subset is not None:
        raise AbcError("'subset' is not implemented")
    if callable(statistic):
        statistic = getattr(statistic, "__name__", "mean")
    eng = engine()
    ss = to_table(sumstat, eng.device)
    tgt = _to_vec(target, eng.device)
    if tgt.numel() != ss.shape[0]:
        raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
    return _cv.gfit(eng, tgt, ss, int(nb_replicate), float(tol), statistic, samples, default_seed if seed is None else seed)
