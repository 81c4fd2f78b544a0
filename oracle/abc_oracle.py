"""CPU oracle for the ABC acceptance-and-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product (``abc_dls_b200``) never imports it and has no CPU path.

**PARITY UNPINNED.**  The arithmetic ABC-DLS runs on this path is not in ``/root/reference``: it is
the third-party CRAN package ``abc`` pinned at 2.2.1 (reference ``requirements_all.yml:22``), reached
through rpy2 3.5.11 (``requirements.yml:8``).  Neither R, rpy2 nor the CRAN tarball is available in
the build container, and the reference's own tests (``tests/test_ABCDLS.py:118-138,183-204,230-242``)
only assert that output files exist.  This file therefore restates the *published* algorithm of
``abc::abc``, ``abc::postpr`` and ``abc::summary.abc`` (order of operations in SURVEY.md appendix A)
and anchors it on the reference call sites:

  * ``src/Classes/ABC.py:1949-1953, 1964``  abc.abc(target, param, sumstat, method, tol)
  * ``src/Classes/ABC.py:919``              abc.postpr(target, index, sumstat, tol, method)
  * ``src/Classes/ABC.py:2806-2840``        abc.abc + summary()[0] / [6]  (SMC min / max)
  * ``src/Classes/ABC.py:2921-2946``        updating_newrange
  * ``src/Classes/ABC.py:2881-2919``        noise_injection_update
  * ``src/Classes/ABC.py:2974-2990``        narrowing_params
  * ``src/Classes/ABC.py:3016-3028``        lmrd_calculation

Every choice that had to be restated from the R package (not checkable here) is a named keyword so it
can be flipped the day a live R is reachable: ``cap_rule``, ``bandwidth``, ``recentre``, ``hcorr``.

Style: plain numpy fp64, written for auditability, one R statement per line where possible.
Base-R helpers are restated with their documented semantics (``median`` even-N rule, ``mean`` in long
double with the second refinement pass, ``mad`` constant 1.4826, ``quantile`` type 7, ``lsfit`` =
Householder QR on sqrt(w)-scaled rows with zero-weight rows dropped).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

MAD_CONSTANT = 1.4826  # stats::mad default `constant`


class AbcError(RuntimeError):
    """Restates an R ``stop()`` raised inside abc / postpr (message text kept)."""


# --------------------------------------------------------------------------------------
# base-R helpers
# --------------------------------------------------------------------------------------
def r_mean(x: np.ndarray) -> float:
    """R ``mean.default`` on a double vector: long-double sum, divide, one refinement pass
    (R ``summary.c: real_mean``)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n == 0:
        return float("nan")
    xl = x.astype(np.longdouble)
    s = xl.sum(dtype=np.longdouble) / np.longdouble(n)
    t = (xl - s).sum(dtype=np.longdouble)
    s = s + t / np.longdouble(n)
    return float(s)


def r_median(x: np.ndarray) -> float:
    """R ``median.default``: odd n -> the middle order statistic; even n -> ``mean`` of the two middle."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    if n == 0:
        return float("nan")
    half = (n + 1) // 2  # 1-based
    if n % 2 == 1:
        return float(np.partition(x, half - 1)[half - 1])
    p = np.partition(x, [half - 1, half])
    return r_mean(p[half - 1: half + 1])


def r_mad(x: np.ndarray) -> float:
    """R ``stats::mad(x)`` = 1.4826 * median(abs(x - median(x)))."""
    x = np.asarray(x, dtype=np.float64).ravel()
    return MAD_CONSTANT * r_median(np.abs(x - r_median(x)))


def r_quantile7(x: np.ndarray, probs: Sequence[float]) -> np.ndarray:
    """R ``quantile(x, probs, type = 7)`` including R's fuzz on the index."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = x.size
    out = []
    for p in probs:
        index = 1 + (n - 1) * p
        lo = math.floor(index + 4 * np.finfo(float).eps)
        hi = math.ceil(index - 4 * np.finfo(float).eps)
        xl, xh = x[lo - 1], x[hi - 1]
        h = index - lo
        # R: x[lo] when h == 0 (avoids Inf*0), else (1-h)*x[lo] + h*x[hi]
        out.append(float(xl) if h == 0 or xl == xh else float((1 - h) * xl + h * xh))
    return np.array(out)


def r_weighted_quantile(x: np.ndarray, w: np.ndarray, taus: Sequence[float]) -> np.ndarray:
    """``quantreg::rq(x ~ 1, tau, weights = w)$coef`` for an intercept-only model.

    The minimiser of sum_i w_i * rho_tau(x_i - q) is an order statistic x_(k) with
    cumw(k-1) <= tau*W <= cumw(k).  ``rq``'s simplex ('br') returns a vertex, i.e. a data point; when
    tau*W hits a knot exactly the solution is not unique and the vertex returned by R is not
    restated here (the lower one is used).  Zero-weight points never bind.
    """
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    order = np.argsort(x, kind="stable")
    xs, ws = x[order], w[order]
    cw = np.cumsum(ws)
    W = cw[-1]
    out = []
    for tau in taus:
        k = int(np.searchsorted(cw, tau * W, side="left"))
        k = min(max(k, 0), x.size - 1)
        out.append(float(xs[k]))
    return np.array(out)


def lsfit(X: np.ndarray, Y: np.ndarray, w: np.ndarray, rank_tol: float = 1e-7) -> np.ndarray:
    """R ``lsfit(X, Y, wt = w)`` coefficients ((d+1) x P, intercept first).

    lsfit multiplies rows by sqrt(wt), drops zero-weight rows, and solves by Householder QR
    (LINPACK dqrls, column rank tolerance 1e-7).  A rank-deficient design is *not* restated
    (R drops the aliased column with a warning); it raises here.
    """
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    if Y.ndim == 1:
        Y = Y[:, None]
    keep = w > 0
    sw = np.sqrt(w[keep])
    A = np.column_stack([np.ones(int(keep.sum())), X[keep]]) * sw[:, None]
    B = Y[keep] * sw[:, None]
    Q, R = np.linalg.qr(A, mode="reduced")
    diag = np.abs(np.diag(R))
    if A.shape[0] < A.shape[1] or diag.min() <= rank_tol * diag.max():
        raise AbcError("lsfit: X matrix deemed to be singular")
    return np.linalg.solve(R, Q.T @ B)


# --------------------------------------------------------------------------------------
# abc::abc  (methods rejection / loclinear, transf = none, kernel = epanechnikov)
# --------------------------------------------------------------------------------------
@dataclass
class AbcOracleResult:
    method: str
    region: np.ndarray  # bool[N]   (R: wt1)
    idx: np.ndarray  # int64[nacc] ascending row numbers (0-based)
    dist: np.ndarray  # float64[N]
    ds: float
    nacc: int
    mad: np.ndarray  # float64[d]   the divisor actually used per column is mad, or 1 when mad == 0
    unadj_values: np.ndarray  # nacc x P
    ss: np.ndarray  # nacc x d   (unscaled)
    weights: np.ndarray  # nacc
    adj_values: Optional[np.ndarray] = None  # nacc x P  (loclinear)
    residuals: Optional[np.ndarray] = None
    coef: Optional[np.ndarray] = None  # (d+1) x P  fit1, uncentred design [1, scaled ss]
    coef_h: Optional[np.ndarray] = None  # (d+1) x P  fit2
    pred: Optional[np.ndarray] = None  # P   prediction at the target (after the_m shift)
    sigma2: Optional[np.ndarray] = None
    aic: Optional[float] = None
    bic: Optional[float] = None
    warnings: List[str] = field(default_factory=list)


def _as_matrix(a) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return a


def scale_and_distance(target: np.ndarray, sumstat: np.ndarray):
    """R abc(): ``normalise`` every column by its MAD (skip when MAD == 0), the target likewise, then
    the Euclidean distance accumulated sequentially over columns.  Returns (scaled, t, mad, dist)."""
    N, d = sumstat.shape
    mad = np.empty(d)
    scaled = np.empty_like(sumstat)
    t = np.empty(d)
    for j in range(d):
        m = r_mad(sumstat[:, j])
        mad[j] = m
        if m == 0:
            scaled[:, j] = sumstat[:, j]
            t[j] = target[j]
        else:
            scaled[:, j] = sumstat[:, j] / m
            t[j] = target[j] / m
    sum1 = np.zeros(N)
    for j in range(d):
        diff = scaled[:, j] - t[j]
        sum1 = sum1 + diff * diff  # R: ^2 on doubles is x*x; no fused multiply-add
    return scaled, t, mad, np.sqrt(sum1)


def tolerance_cut(dist: np.ndarray, tol: float, cap_rule: str = "first_nacc_in_row_order"):
    """R abc(): nacc <- ceiling(N*tol); ds <- sort(dist)[nacc]; wt1 <- dist <= ds;
    wt1 <- wt1 & cumsum(wt1) <= nacc."""
    N = dist.size
    nacc = int(math.ceil(N * tol))
    if nacc < 1 or nacc > N:
        raise AbcError("tol must give 1 <= ceiling(N*tol) <= N")
    ds = float(np.partition(dist, nacc - 1)[nacc - 1])
    wt1 = dist <= ds
    if cap_rule == "first_nacc_in_row_order":
        wt1 = wt1 & (np.cumsum(wt1) <= nacc)
    elif cap_rule != "none":
        raise ValueError(cap_rule)
    return nacc, ds, wt1


def abc(target, param, sumstat, tol: float, method: str, hcorr: bool = True,
        kernel: str = "epanechnikov", cap_rule: str = "first_nacc_in_row_order",
        bandwidth: str = "ds", recentre: str = "unweighted_mean") -> AbcOracleResult:
    """Restatement of ``abc::abc`` 2.2.1 for method in {rejection, loclinear}."""
    if method not in ("rejection", "loclinear"):
        raise AbcError("oracle covers methods rejection and loclinear only")
    if kernel != "epanechnikov":
        raise AbcError("oracle covers the epanechnikov kernel only")
    param = _as_matrix(param)
    sumstat = _as_matrix(sumstat)
    target = np.asarray(target, dtype=np.float64).ravel()
    N, d = sumstat.shape
    P = param.shape[1]
    if target.size != d:
        raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
    if param.shape[0] != N:
        raise AbcError("'param' and 'sumstat' must have the same number of rows")
    if not (np.isfinite(sumstat).all() and np.isfinite(param).all() and np.isfinite(target).all()):
        raise AbcError("non-finite input (NA rows / subset are not restated)")
    # stop if zero var in sumstat:  cond1 <- !any(apply(sumstat, 2, function(x) length(unique(x))-1))
    if not any(np.unique(sumstat[:, j]).size > 1 for j in range(d)):
        raise AbcError("Zero variance in the summary statistics.")

    scaled, t, mad, dist = scale_and_distance(target, sumstat)
    nacc, ds, wt1 = tolerance_cut(dist, tol, cap_rule)
    idx = np.flatnonzero(wt1)
    unadj = param[wt1, :]
    ss = sumstat[wt1, :]
    res = AbcOracleResult(method=method, region=wt1, idx=idx, dist=dist, ds=ds, nacc=nacc, mad=mad,
                          unadj_values=unadj, ss=ss, weights=np.ones(idx.size))

    statvar = np.array([np.unique(ss[:, j]).size > 1 for j in range(d)])
    cond2 = not statvar.any()
    if method == "rejection":
        if cond2:
            res.warnings.append("Zero variance in the summary statistics in the selected region. "
                                "Check summary statistics, consider larger tolerance.")
        return res
    if cond2:
        raise AbcError("Zero variance in the summary statistics in the selected region. Try: checking "
                       "summary statistics, choosing larger tolerance, or rejection method.")

    h = ds if bandwidth == "ds" else float(bandwidth)
    q = dist[wt1] / h
    w = 1.0 - q * q  # epanechnikov
    res.weights = w

    X = scaled[wt1, :]
    coef = lsfit(X, unadj, w)  # (d+1) x P
    design = np.column_stack([np.ones(idx.size), X])
    pred = np.concatenate([[1.0], t]) @ coef  # P
    residuals = unadj - design @ coef
    if recentre == "unweighted_mean":
        the_m = np.array([r_mean(residuals[:, p]) for p in range(P)])
    elif recentre == "none":
        the_m = np.zeros(P)
    else:
        raise ValueError(recentre)
    residuals = residuals - the_m
    pred = pred + the_m
    sw = w.sum()
    sigma2 = (residuals * residuals * w[:, None]).sum(axis=0) / sw
    res.coef, res.sigma2 = coef, sigma2
    with np.errstate(divide="ignore"):
        res.aic = float(idx.size * np.log(sigma2).sum() + 2 * (d + 1) * P)
        res.bic = float(idx.size * np.log(sigma2).sum() + math.log(idx.size) * (d + 1) * P)
    if hcorr:
        with np.errstate(divide="ignore"):
            z = np.log(residuals * residuals)
        coef_h = lsfit(X, z, w)
        pred_sd = np.sqrt(np.exp(np.concatenate([[1.0], t]) @ coef_h))  # P
        pred_si = np.sqrt(np.exp(design @ coef_h))  # nacc x P
        residuals = pred_sd * residuals / pred_si
        res.coef_h = coef_h
    res.adj_values = pred + residuals
    res.residuals = residuals
    res.pred = pred
    return res


# --------------------------------------------------------------------------------------
# abc::postpr (rejection)
# --------------------------------------------------------------------------------------
@dataclass
class PostprOracleResult:
    models: List[str]
    counts: np.ndarray  # int64[K] accepted simulations per model, ``models`` order (sorted levels)
    pred: np.ndarray  # counts / nacc
    values: np.ndarray  # accepted model labels, row order
    idx: np.ndarray
    ds: float
    nacc: int
    dist: np.ndarray
    mad: np.ndarray

    def bayes_factors(self) -> np.ndarray:
        with np.errstate(divide="ignore", invalid="ignore"):
            return self.pred[:, None] / self.pred[None, :]


def postpr(target, index, sumstat, tol: float, method: str = "rejection",
           cap_rule: str = "first_nacc_in_row_order") -> PostprOracleResult:
    """Restatement of ``abc::postpr(method = "rejection")``: same scaling / distance / cut as abc(),
    then the proportion of each model among the accepted rows."""
    if method != "rejection":
        raise AbcError("oracle covers postpr(method='rejection') only")
    sumstat = _as_matrix(sumstat)
    target = np.asarray(target, dtype=np.float64).ravel()
    index = np.asarray(index).astype(str)
    N, d = sumstat.shape
    if index.size != N:
        raise AbcError("'index' must be the same length as the number of rows in 'sumstat'.")
    if target.size != d:
        raise AbcError("Number of summary statistics in 'target' has to be the same as in 'sumstat'.")
    models = sorted(set(index.tolist()))  # levels(factor(index))
    if len(models) == 1:
        raise AbcError("At least two different models must be given.")
    if not any(np.unique(sumstat[:, j]).size > 1 for j in range(d)):
        raise AbcError("Zero variance in the summary statistics.")
    _, _, mad, dist = scale_and_distance(target, sumstat)
    nacc, ds, wt1 = tolerance_cut(dist, tol, cap_rule)
    values = index[wt1]
    counts = np.array([(values == m).sum() for m in models], dtype=np.int64)
    return PostprOracleResult(models=models, counts=counts, pred=counts / values.size, values=values,
                              idx=np.flatnonzero(wt1), ds=ds, nacc=nacc, dist=dist, mad=mad)


# --------------------------------------------------------------------------------------
# summary.abc  (7 x P table: Min, 2.5%, Median, Mean, Mode, 97.5%, Max)
# --------------------------------------------------------------------------------------
SUMMARY_ROWS = ("Min.:", "2.5% Perc.:", "Median:", "Mean:", "Mode:", "97.5% Perc.:", "Max.:")


def r_bw_nrd0(x: np.ndarray) -> float:
    """stats::bw.nrd0 (R 4.3): 0.9 * min(sd, IQR / 1.34) * n^-0.2, with R's fallbacks when that is 0."""
    x = np.asarray(x, dtype=np.float64)
    hi = float(np.std(x, ddof=1))
    q = r_quantile7(x, [0.25, 0.75])
    lo = min(hi, float(q[1] - q[0]) / 1.34)
    if not lo:
        lo = hi or abs(float(x[0])) or 1.0
    return 0.9 * lo * x.size ** (-0.2)


def r_density(x: np.ndarray, weights=None, n: int = 512, cut: float = 3.0):
    """stats::density.default (R 4.3, gaussian kernel, bw = "nrd0", old grid coordinates): linear binning of the
    (weighted) sample onto 2n points over [lo, up] = [min - 7 bw, max + 7 bw], circular convolution with the normal
    kernel by FFT, clamping at 0, linear interpolation onto n points over [min - 3 bw, max + 3 bw]."""
    x = np.asarray(x, dtype=np.float64)
    nx = x.size
    w = np.full(nx, 1.0 / nx) if weights is None else np.asarray(weights, dtype=np.float64)
    bw = r_bw_nrd0(x)
    frm, to = x.min() - cut * bw, x.max() + cut * bw
    lo, up = frm - 4.0 * bw, to + 4.0 * bw
    # C_BinDist
    y = np.zeros(2 * n)
    xdelta = (up - lo) / (n - 1)
    xpos = (x - lo) / xdelta
    ix = np.floor(xpos).astype(np.int64)
    fx = xpos - ix
    inside = (ix >= 0) & (ix <= n - 2)
    np.add.at(y, ix[inside], w[inside] * (1 - fx[inside]))
    np.add.at(y, ix[inside] + 1, w[inside] * fx[inside])
    left = ix == -1
    np.add.at(y, np.zeros(int(left.sum()), dtype=np.int64), w[left] * fx[left])
    right = ix == n - 1
    np.add.at(y, ix[right], w[right] * (1 - fx[right]))
    kords = np.linspace(0.0, 2.0 * (up - lo), 2 * n)
    kords[n + 1:] = -kords[n - 1:0:-1]
    kords = np.exp(-0.5 * (kords / bw) ** 2) / (bw * np.sqrt(2.0 * np.pi))
    conv = np.fft.ifft(np.fft.fft(y) * np.conj(np.fft.fft(kords))) * (2 * n)    # R's inverse fft is unnormalised
    dens = np.maximum(0.0, conv.real[:n] / (2 * n))
    xords = np.linspace(lo, up, n)
    xout = np.linspace(frm, to, n)
    return xout, np.interp(xout, xords, dens)


def r_density_mode(x: np.ndarray, weights=None) -> float:
    """abc:::getmode: the grid point of ``density(x, weights = w / sum(w))`` with the largest density (first one)."""
    if weights is not None:
        weights = np.asarray(weights, dtype=np.float64)
        weights = weights / weights.sum()
    gx, gy = r_density(x, weights)
    return float(gx[int(np.argmax(gy))])


def summary_abc(res: AbcOracleResult, intvl: float = 0.95) -> np.ndarray:
    """``summary.abc``: rejection -> plain statistics of unadj.values; loclinear -> min/max plain,
    weighted mean, weighted quantiles (rq), mode = argmax of the (weighted) kernel density."""
    a = (1 - intvl) / 2
    if res.method == "rejection":
        x = res.unadj_values
        out = np.full((7, x.shape[1]), np.nan)
        for p in range(x.shape[1]):
            q = r_quantile7(x[:, p], [a, 0.5, 1 - a])
            out[:, p] = [x[:, p].min(), q[0], q[1], r_mean(x[:, p]), r_density_mode(x[:, p]), q[2], x[:, p].max()]
        return out
    x, w = res.adj_values, res.weights
    out = np.full((7, x.shape[1]), np.nan)
    for p in range(x.shape[1]):
        q = r_weighted_quantile(x[:, p], w, [a, 0.5, 1 - a])
        out[:, p] = [x[:, p].min(), q[0], q[1], float((x[:, p] * w).sum() / w.sum()), r_density_mode(x[:, p], w),
                     q[2], x[:, p].max()]
    return out


def r_format_real(x, digits: int = 4) -> List[str]:
    """R ``formatReal`` as print.default applies it to one column of a numeric matrix (src/main/format.c, restated):
    per element the decimal exponent e and the number of significant digits sig (<= digits, trailing zeros dropped);
    fixed notation needs max(sig - e - 1, 0) decimals and max(e + 1, 1) digits before the point; the column is printed in
    fixed notation with the largest decimal count unless scientific notation is narrower (R_print.scipen = 0)."""
    x = np.asarray(x, dtype=np.float64)
    es, sigs = [], []
    for v in x:
        if v == 0:
            es.append(0), sigs.append(1)
            continue
        e = int(math.floor(math.log10(abs(v))))
        m = round(abs(v) / 10.0 ** e, digits - 1)
        if m >= 10.0:
            m, e = m / 10.0, e + 1
        q = int(round(m * 10 ** (digits - 1)))
        sig = digits
        while sig > 1 and q % 10 == 0:
            q //= 10
            sig -= 1
        es.append(e), sigs.append(sig)
    neg = bool((x < 0).any())
    rgt = max(max(s_ - e - 1, 0) for s_, e in zip(sigs, es))
    left = max(max(e + 1, 1) + (1 if v < 0 else 0) for v, e in zip(x, es))
    w_fixed = left + rgt + (1 if rgt else 0)
    mxe = max(sigs) - 1
    w_sci = int(neg) + 1 + (mxe + 1 if mxe else 0) + 4
    if w_fixed <= w_sci:
        return ["%.*f" % (rgt, v) for v in x]
    return ["%.*e" % (mxe, v) for v in x]


# --------------------------------------------------------------------------------------
# ABC-DLS SMC range logic (pandas in the reference; numpy here, same arithmetic order)
# --------------------------------------------------------------------------------------
def updating_newrange(new_min, new_max, old_min, old_max, decrease: float = 0.95, ndigits: int = 3):
    """ABC.py:2921-2946.  Returns (min, max, decrease) rounded to ``ndigits``."""
    new_min = np.array(new_min, dtype=np.float64)
    new_max = np.array(new_max, dtype=np.float64)
    old_min = np.asarray(old_min, dtype=np.float64)
    old_max = np.asarray(old_max, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        dec = (new_max - new_min) / (old_max - old_min)
        weak = (dec > decrease) & (dec < 1)
        new_min[weak], new_max[weak] = old_min[weak], old_max[weak]
        zero = dec == 0
        # newrange.loc[dec == 0] = oldrange.loc[...]: min/max come from oldrange; the 'decrease' cell
        # is overwritten on the next line anyway
        new_min[zero], new_max[zero] = old_min[zero], old_max[zero]
        dec = (new_max - new_min) / (old_max - old_min)
    return np.round(new_min, ndigits), np.round(new_max, ndigits), np.round(dec, ndigits)


def noise_injection_update(new_min, new_max, dec, old_min, old_max, increase: float = 0.005,
                           hard_min=None, hard_max=None, decrease: float = 0.95, ndigits: int = 3):
    """ABC.py:2881-2919."""
    new_min = np.array(new_min, dtype=np.float64)
    new_max = np.array(new_max, dtype=np.float64)
    dec = np.array(dec, dtype=np.float64)
    old_min = np.asarray(old_min, dtype=np.float64)
    old_max = np.asarray(old_max, dtype=np.float64)
    dist = (new_max - new_min) * increase * 0.5
    sel = dec > decrease
    lo = (new_min - dist)[sel]
    hi = (new_max + dist)[sel]
    new_min[sel] = lo
    new_max[sel] = hi
    if hard_min is not None:
        new_min = np.maximum(np.asarray(hard_min, dtype=np.float64), new_min)
        new_max = np.minimum(np.asarray(hard_max, dtype=np.float64), new_max)
    with np.errstate(divide="ignore", invalid="ignore"):
        dec = (new_max - new_min) / (old_max - old_min)
    still_zero = np.round(new_max, ndigits) == np.round(new_min, ndigits)
    if still_zero.any():
        new_min[still_zero] = new_min[still_zero] - 10 ** (-ndigits)
        dec[still_zero] = 1.0
    return np.round(new_min, ndigits), np.round(new_max, ndigits), np.round(dec, ndigits)


def narrowing_params(params: np.ndarray, parmin, parmax) -> np.ndarray:
    """ABC.py:2974-2990: rows with every parameter inside [min, max] (inclusive) -> csv line
    numbers (0-based row + 2)."""
    params = _as_matrix(params)
    keep = np.ones(params.shape[0], dtype=bool)
    for j in range(params.shape[1]):
        keep &= (params[:, j] >= parmin[j]) & (params[:, j] <= parmax[j])
    return np.flatnonzero(keep) + 2


def lmrd_calculation(new_min, new_max, hard_min, hard_max) -> float:
    """ABC.py:3016-3028: log(mean(|new range| / |hard range|))."""
    nd = np.abs(np.asarray(new_max, dtype=np.float64) - np.asarray(new_min, dtype=np.float64))
    hd = np.abs(np.asarray(hard_max, dtype=np.float64) - np.asarray(hard_min, dtype=np.float64))
    return float(np.log((nd / hd).mean()))


def smc_round(target, param, sumstat, tol, method, param_all, decrease=0.95, increase=0.0,
              hard_min=None, hard_max=None) -> Dict[str, np.ndarray]:
    """One ABC-DLS SMC after-train round (ABC.py:2709-2731) in per-column mode: P separate 1-D abc()
    calls -> posterior min/max -> updating_newrange -> [noise_injection_update] -> narrowing."""
    param = _as_matrix(param)
    sumstat = _as_matrix(sumstat)
    target = np.asarray(target, dtype=np.float64).ravel()
    P = param.shape[1]
    pmin, pmax = np.empty(P), np.empty(P)
    for c in range(P):
        r = abc(target[c:c + 1], param[:, c], sumstat[:, c], tol, method)
        tab = summary_abc(r)
        pmin[c], pmax[c] = tab[0, 0], tab[6, 0]
    param_all = _as_matrix(param_all)
    old_min, old_max = param_all.min(axis=0), param_all.max(axis=0)
    nmin, nmax, dec = updating_newrange(pmin, pmax, old_min, old_max, decrease)
    out = {"post_min": pmin, "post_max": pmax, "old_min": old_min, "old_max": old_max}
    if increase > 0:
        nmin, nmax, dec = noise_injection_update(nmin, nmax, dec, old_min, old_max, increase, hard_min,
                                                 hard_max, decrease)
        if hard_min is not None:
            out["lmrd"] = lmrd_calculation(nmin, nmax, hard_min, hard_max)
    out.update({"min": nmin, "max": nmax, "decrease": dec,
                "lines": narrowing_params(param_all, nmin, nmax)})
    return out


# --------------------------------------------------------------------------------------
# Leave-one-out tools of the R package: abc::cv4abc, abc::cv4postpr, abc::gfit and their summaries.
# Reference call sites: ABC.py:1900-1902, 1913 (cv4abc + summary()[0][0] / the 1 x P table), ABC.py:886-892 (cv4postpr +
# summary), ABC.py:961-963 (gfit + summary).  Restated from the published R sources (abc 2.2.1: cv4abc.R, cv4postpr.R,
# gfit.R), unpinned like the rest.  R draws the held-out rows with sample() (Mersenne twister + R's rejection sampling):
# that stream is not restated; the rows are an explicit argument (``cvsamples``) and default to a numpy draw.
# --------------------------------------------------------------------------------------
_CV_ROW = {"median": 2, "mean": 3, "mode": 4}


@dataclass
class Cv4abcOracleResult:
    cvsamples: np.ndarray  # int64[nval] 0-based held-out rows
    tols: np.ndarray  # sorted ascending (R: tols <- sort(tols))
    true: np.ndarray  # nval x P   param[cvsamples, ]
    estim: Dict[float, np.ndarray]  # tol -> nval x P point estimates


def cv4abc(param, sumstat, nval: int, tols, method: str, statistic: str = "median", hcorr: bool = True,
           cvsamples=None, rng=None) -> Cv4abcOracleResult:
    """``abc::cv4abc``: for every tolerance and every held-out row i: abc(target = sumstat[i, ], param = param[-i, ],
    sumstat = sumstat[-i, ]) -- the row is physically removed, so N - 1 rows enter the MADs and nacc =
    ceiling((N - 1) * tol) -- and the point estimate is row ``statistic`` of summary.abc."""
    param, sumstat = _as_matrix(param), _as_matrix(sumstat)
    N = sumstat.shape[0]
    if cvsamples is None:
        rng = rng if rng is not None else np.random.default_rng()
        cvsamples = rng.choice(N, size=nval, replace=False)       # sample(1:numsim, nval)
    cvsamples = np.asarray(cvsamples, dtype=np.int64)
    tols = np.sort(np.atleast_1d(np.asarray(tols, dtype=np.float64)))
    estim: Dict[float, np.ndarray] = {}
    for tol in tols:
        res = np.empty((cvsamples.size, param.shape[1]))
        for i, s in enumerate(cvsamples):
            keep = np.ones(N, dtype=bool)
            keep[s] = False
            sub = abc(sumstat[s], param[keep], sumstat[keep], float(tol), method, hcorr=hcorr)
            res[i] = summary_abc(sub)[_CV_ROW[statistic]]
        estim[float(tol)] = res
    return Cv4abcOracleResult(cvsamples=cvsamples, tols=tols, true=param[cvsamples], estim=estim)


def summary_cv4abc(cv: Cv4abcOracleResult) -> np.ndarray:
    """``summary.cv4abc``: prediction error, one row per tolerance: colSums((estim - true)^2) / (var(true) * nval)."""
    nval = cv.true.shape[0]
    truevar = cv.true.var(axis=0, ddof=1) * nval
    return np.stack([((cv.estim[float(t)] - cv.true) ** 2).sum(axis=0) / truevar for t in cv.tols])


@dataclass
class Cv4postprOracleResult:
    cvsamples: np.ndarray  # nval rows per model, models in sorted-level order (R: unlist(tapply(..., sample, nval)))
    tols: np.ndarray
    true: np.ndarray  # model label of every held-out row
    models: List[str]
    estim: Dict[float, np.ndarray]  # tol -> (nval * K) x K posterior model probabilities


def cv4postpr(index, sumstat, nval: int, tols, method: str = "rejection", cvsamples=None, rng=None) -> Cv4postprOracleResult:
    """``abc::cv4postpr``: ``nval`` held-out rows PER MODEL; each is classified with postpr() on the remaining N - 1
    rows (physically removed) and the posterior model probabilities are recorded."""
    sumstat = _as_matrix(sumstat)
    index = np.asarray(index).astype(str)
    N = sumstat.shape[0]
    models = sorted(set(index.tolist()))
    if cvsamples is None:
        rng = rng if rng is not None else np.random.default_rng()
        cvsamples = np.concatenate([rng.choice(np.flatnonzero(index == m), size=nval, replace=False) for m in models])
    cvsamples = np.asarray(cvsamples, dtype=np.int64)
    tols = np.sort(np.atleast_1d(np.asarray(tols, dtype=np.float64)))
    estim: Dict[float, np.ndarray] = {}
    for tol in tols:
        res = np.zeros((cvsamples.size, len(models)))
        for i, s in enumerate(cvsamples):
            keep = np.ones(N, dtype=bool)
            keep[s] = False
            sub = postpr(sumstat[s], index[keep], sumstat[keep], float(tol), method)
            for m, p in zip(sub.models, sub.pred):            # a model whose last row was held out may be missing
                res[i, models.index(m)] = p
        estim[float(tol)] = res
    return Cv4postprOracleResult(cvsamples=cvsamples, tols=tols, true=index[cvsamples], models=models, estim=estim)


def summary_cv4postpr(cv: Cv4postprOracleResult):
    """``summary.cv4postpr``: per tolerance the confusion matrix table(true, argmax posterior) (ties: first model, R's
    which.max) and the mean posterior model probabilities per true model."""
    K = len(cv.models)
    ti = np.array([cv.models.index(m) for m in cv.true])
    conf, probs = {}, {}
    for t in cv.tols:
        e = cv.estim[float(t)]
        pm = e.argmax(axis=1)
        cm = np.zeros((K, K), dtype=np.int64)
        np.add.at(cm, (ti, pm), 1)
        conf[float(t)] = cm
        probs[float(t)] = np.stack([e[ti == k].mean(axis=0) for k in range(K)])
    return {"conf.matrix": conf, "probs": probs}


@dataclass
class GfitOracleResult:
    dist_obs: float
    dist_sim: np.ndarray


def gfit(target, sumstat, nb_replicate: int, tol: float, statistic: str = "mean", samples=None, rng=None) -> GfitOracleResult:
    """``abc::gfit``: goodness-of-fit statistic = ``statistic`` of the distances of the accepted simulations
    (abc(method = "rejection")); its null distribution treats ``nb_replicate`` simulations in turn as the observation,
    each against the table with that row removed."""
    sumstat = _as_matrix(sumstat)
    target = np.asarray(target, dtype=np.float64).ravel()
    N = sumstat.shape[0]
    stat = {"mean": r_mean, "median": r_median}[statistic]
    if samples is None:
        rng = rng if rng is not None else np.random.default_rng()
        samples = rng.choice(N, size=nb_replicate, replace=False)
    samples = np.asarray(samples, dtype=np.int64)
    obs = abc(target, sumstat[:, 0], sumstat, tol, "rejection")
    dist_sim = np.empty(samples.size)
    for i, s in enumerate(samples):
        keep = np.ones(N, dtype=bool)
        keep[s] = False
        r = abc(sumstat[s], sumstat[keep, 0], sumstat[keep], tol, "rejection")
        dist_sim[i] = stat(r.dist[r.region])
    return GfitOracleResult(dist_obs=stat(obs.dist[obs.region]), dist_sim=dist_sim)


def summary_gfit(g: GfitOracleResult):
    """``summary.gfit``: p-value = share of null statistics above the observed one, summary() of the null."""
    q = r_quantile7(g.dist_sim, [0.0, 0.25, 0.5, 0.75, 1.0])
    return {"pvalue": float((g.dist_sim > g.dist_obs).sum() / g.dist_sim.size),
            "s.dist.sim": np.array([q[0], q[1], q[2], r_mean(g.dist_sim), q[3], q[4]]),   # Min 1stQu Median Mean 3rdQu Max
            "dist.obs": g.dist_obs}
