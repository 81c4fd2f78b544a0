"""``summary.abc`` on the posterior sample (7 x P table: Min, 2.5%, Median, Mean, Mode, 97.5%, Max).

Min / Max / (weighted) Mean come from the ``k_col_stats`` kernel (they are what ABC-DLS's SMC reads,
ABC.py:2833, 2837).  The percentile rows are SURVEY §8f "next" work: they are computed here with a
device sort (R ``quantile`` type 7 for rejection; weighted order statistic = ``quantreg::rq(x ~ 1)`` for
loclinear).  The Mode row is ``abc:::getmode``: the grid point of R's ``density()`` (gaussian kernel, ``bw.nrd0``, linear
binning onto 1024 points + FFT convolution, 512 output points) with the largest (weighted) density.
"""
from __future__ import annotations

import numpy as np
import torch

ROWS = ("Min.:", "2.5% Perc.:", "Median:", "Mean:", "Mode:", "97.5% Perc.:", "Max.:")


class SummaryTable(np.ndarray):
    """7 x P float64 table.  Integer indexing follows R / rpy2 (column-major flat): for a single
    parameter ``tab[0]`` is Min and ``tab[6]`` is Max, as ABC-DLS uses it (ABC.py:2833, 2837);
    ``numpy.array(tab)`` gives the plain 7 x P matrix (ABC.py:2854, 2858)."""

    def __new__(cls, a, names=None):
        obj = np.asarray(a, dtype=np.float64).view(cls)
        obj.names = list(names) if names is not None else None
        return obj

    def __array_finalize__(self, obj):
        self.names = getattr(obj, "names", None)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)) and self.ndim == 2:
            return float(np.asarray(self).flatten(order="F")[key])
        return np.asarray(self)[key]


def _quantile7(xs: torch.Tensor, p: float) -> torch.Tensor:
    n = xs.shape[0]
    index = 1 + (n - 1) * p
    eps4 = 4 * np.finfo(float).eps
    lo, hi = int(np.floor(index + eps4)), int(np.ceil(index - eps4))
    h = index - lo
    xl, xh = xs[lo - 1], xs[hi - 1]
    if h == 0:
        return xl
    return torch.where(xl == xh, xl, (1 - h) * xl + h * xh)


def _density_mode(xs: torch.Tensor, ws, n_grid: int = 512, cut: float = 3.0) -> torch.Tensor:
    """Mode of R's ``density(x, weights = w / sum(w))`` (stats::density.default of R 4.3).  xs: the sample SORTED
    ascending, ws: its weights in the same order (None = equal weights).  Deterministic: the linear binning is done
    with prefix sums over the sorted sample instead of atomic adds."""
    n = xs.numel()
    dev, f64 = xs.device, torch.float64
    sd = xs.std(unbiased=True)
    iqr = _quantile7(xs, 0.75) - _quantile7(xs, 0.25)
    lo_ = torch.minimum(sd, iqr / 1.34)
    lo_ = torch.where(lo_ == 0, torch.where(sd == 0, torch.where(xs[0] == 0, torch.ones_like(sd), xs[0].abs()), sd), lo_)
    bw = 0.9 * lo_ * n ** (-0.2)                                   # bw.nrd0 (of the UNWEIGHTED sample, like R)
    frm, to = xs[0] - cut * bw, xs[-1] + cut * bw
    lo, up = frm - 4.0 * bw, to + 4.0 * bw
    w = torch.full((n,), 1.0 / n, dtype=f64, device=dev) if ws is None else ws / ws.sum()
    N = n_grid
    xdelta = (up - lo) / (N - 1)
    xpos = (xs - lo) / xdelta
    ix = torch.floor(xpos)
    fx = xpos - ix
    ix = ix.to(torch.int64).clamp_(0, N - 2)                        # the data lies 7 bw inside [lo, up]
    # C_BinDist: y[ix] += w (1 - fx); y[ix + 1] += w fx   -- ix is non-decreasing, so segment sums = prefix differences
    c0 = torch.cat([torch.zeros(1, dtype=f64, device=dev), torch.cumsum(w * (1 - fx), 0)])
    c1 = torch.cat([torch.zeros(1, dtype=f64, device=dev), torch.cumsum(w * fx, 0)])
    edges = torch.searchsorted(ix, torch.arange(N + 1, device=dev, dtype=torch.int64), right=False)
    y = torch.zeros(2 * N, dtype=f64, device=dev)
    y[:N] += c0[edges[1:]] - c0[edges[:-1]]
    y[1:N + 1] += c1[edges[1:]] - c1[edges[:-1]]
    k = torch.linspace(0.0, 1.0, 2 * N, dtype=f64, device=dev) * (2.0 * (up - lo))
    k = torch.cat([k[:N + 1], -k[1:N].flip(0)])
    k = torch.exp(-0.5 * (k / bw) ** 2) / (bw * (2.0 * np.pi) ** 0.5)
    dens = torch.fft.ifft(torch.fft.fft(y) * torch.conj(torch.fft.fft(k))).real[:N].clamp_min(0.0)
    xout = torch.arange(N, dtype=f64, device=dev) * ((to - frm) / (N - 1)) + frm     # numpy / R seq(): i * step + from
    t = (xout - lo) / xdelta
    i0 = torch.floor(t).to(torch.int64).clamp_(0, N - 2)
    fr = t - i0
    yout = dens[i0] * (1 - fr) + dens[i0 + 1] * fr
    return xout[torch.argmax(yout)]


def summary_table_device(eng, values: torch.Tensor, weights, stats: torch.Tensor, weighted: bool, intvl: float = 0.95,
                         names=None):
    """The same 7 x P table from this repo's kernels only (csrc/summary.cu, csrc/select.cu): no sort of the sample.
    Percentiles: radix select (type-7 ranks for rejection, weight histograms for loclinear); Mode: bw.nrd0 from the
    column moments and two selected quartiles, then binning + direct gaussian convolution + argmax on the device."""
    k = eng.k
    n, P = values.shape
    vals = values.t()
    if vals.stride(1) != 1:
        vals = vals.contiguous()                                 # (P, n) column-major, as the engine produced it
    a = (1 - intvl) / 2
    f64 = torch.float64
    out = torch.full((7, P), float("nan"), dtype=f64, device=values.device)
    out[0], out[6] = stats[:, 0], stats[:, 1]

    def q7(p):                                                   # R quantile type 7 from two selected order statistics
        index = 1 + (n - 1) * p
        eps4 = 4 * np.finfo(float).eps
        lo, hi = int(np.floor(index + eps4)), int(np.ceil(index - eps4))
        h = index - lo
        r = eng._select(vals, n, None, lo - 1, hi - 1, local=True)
        xl, xh = r[:, 0], (r[:, 1] if hi != lo else r[:, 0])
        return xl if h == 0 else torch.where(xl == xh, xl, (1 - h) * xl + h * xh)

    if weighted:
        out[3] = stats[:, 3] / stats[:, 5]
        w = weights.contiguous()
        taus = torch.tensor([a, 0.5, 1 - a], dtype=f64, device=values.device)
        q = k.wselect(vals, w, n, P, taus)
        out[1], out[2], out[5] = q[:, 0], q[:, 1], q[:, 2]
    else:
        out[3] = stats[:, 2] / n
        w = None
        out[1], out[2], out[5] = q7(a), q7(0.5), q7(1 - a)
    if n >= 2:
        mv = k.col_moments(vals, n, P)
        sd = torch.sqrt(mv[:, 1] / (n - 1))
        iqr = q7(0.75) - q7(0.25)
        lo_ = torch.minimum(sd, iqr / 1.34)
        x0 = stats[:, 0]
        lo_ = torch.where(lo_ == 0, torch.where(sd == 0, torch.where(x0 == 0, torch.ones_like(sd), x0.abs()), sd), lo_)
        bw = 0.9 * lo_ * n ** (-0.2)                             # bw.nrd0 (of the UNWEIGHTED sample, like R)
        frm, to = stats[:, 0] - 3.0 * bw, stats[:, 1] + 3.0 * bw
        lo, up = frm - 4.0 * bw, to + 4.0 * bw
        out[4] = k.density_mode(vals, w, n, P, lo.contiguous(), up.contiguous(), frm.contiguous(), to.contiguous(),
                                bw.contiguous(), stats[:, 5].contiguous() if weighted else None)
    return SummaryTable(out.cpu().numpy(), names)


def summary_table(values: torch.Tensor, weights, stats: torch.Tensor, weighted: bool, intvl: float = 0.95, names=None,
                  eng=None):
    """values (n, P) global posterior sample, weights (n,) or None, stats (P, 6) from k_col_stats.  With an engine whose
    stages are the CUDA library the table comes from this repo's kernels (summary_table_device); the framework-op
    formulation below serves the CPU stage double of the tests."""
    if eng is not None and values.is_cuda and hasattr(eng.k, "wselect"):
        return summary_table_device(eng, values, weights, stats, weighted, intvl, names)
    a = (1 - intvl) / 2
    n, P = values.shape
    out = torch.full((7, P), float("nan"), dtype=torch.float64, device=values.device)
    out[0], out[6] = stats[:, 0], stats[:, 1]
    if weighted:
        out[3] = stats[:, 3] / stats[:, 5]
        for p in range(P):
            xs, order = torch.sort(values[:, p], stable=True)
            ws = weights[order]
            cw = torch.cumsum(ws, 0)
            W = cw[-1]
            if n >= 2:
                out[4, p] = _density_mode(xs, ws)
            for row, tau in ((1, a), (2, 0.5), (5, 1 - a)):
                kk = torch.searchsorted(cw, (tau * W).reshape(1), right=False).clamp(max=n - 1)
                out[row, p] = xs[kk[0]]
    else:
        out[3] = stats[:, 2] / n
        for p in range(P):
            xs, _ = torch.sort(values[:, p])
            out[1, p], out[2, p], out[5, p] = _quantile7(xs, a), _quantile7(xs, 0.5), _quantile7(xs, 1 - a)
            if n >= 2:
                out[4, p] = _density_mode(xs, None)
    return SummaryTable(out.cpu().numpy(), names)


def point_estimate(values: torch.Tensor, weights, stats: torch.Tensor, weighted: bool, statistic: str) -> torch.Tensor:
    """Row ``statistic`` (median | mean | mode) of the summary.abc table only, as a (P,) device tensor: what cv4abc keeps
    of every leave-one-out fit (abc::cv4abc: ``summary.abc(subres)[3 | 4 | 5, ]``).  All parameters in one sort."""
    n, P = values.shape
    if statistic == "mean":
        return stats[:, 3] / stats[:, 5] if weighted else stats[:, 2] / n
    if statistic == "median":
        if not weighted:
            return _quantile7(torch.sort(values, dim=0).values, 0.5)
        xs, order = torch.sort(values, dim=0, stable=True)
        cw = torch.cumsum(weights[order], dim=0).t().contiguous()                # (P, n)
        kk = torch.searchsorted(cw, 0.5 * cw[:, -1:], right=False).clamp(max=n - 1)
        return xs.gather(0, kk.t())[0]
    if statistic == "mode":
        out = torch.full((P,), float("nan"), dtype=torch.float64, device=values.device)
        for p in range(P):
            xs, order = torch.sort(values[:, p], stable=True)
            if n >= 2:
                out[p] = _density_mode(xs, weights[order] if weighted else None)
        return out
    raise ValueError(statistic)


def _sci_parts(x: float, digits: int):
    """(exponent kp, significant digits nsig <= digits after dropping trailing zeros) of x rounded to `digits`
    significant digits - R's scientific() in format.c."""
    if x == 0.0:
        return 0, 1
    mant, exp = f"{abs(x):.{digits - 1}e}".split("e")
    return int(exp), max(1, len(mant.replace(".", "").rstrip("0")))


def format_r_column(values, digits: int = 4):
    """R's formatReal() for one column of a printed numeric matrix: ONE format for the whole column - fixed notation
    with as many decimals as the element that needs most of them to show `digits` significant digits, or scientific
    notation when that would be wider.  Returns the list of cell strings (not yet padded)."""
    vals = [float(v) for v in values]
    fin = [v for v in vals if np.isfinite(v)]
    if not fin:
        return ["NaN" if np.isnan(v) else ("Inf" if v > 0 else "-Inf") for v in vals]
    parts = [_sci_parts(v, digits) for v in fin]
    neg = 1 if any(v < 0 for v in fin) else 0
    rgt = max(max(0, ns - kp - 1) for kp, ns in parts)
    left = max((kp + 1 if kp >= 0 else 1) + (1 if v < 0 else 0) for v, (kp, ns) in zip(fin, parts))
    w_fixed = left + rgt + (1 if rgt > 0 else 0)
    mxe = max(ns for _, ns in parts) - 1
    wexp = 2 if all(abs(kp) < 100 for kp, _ in parts) else 3
    w_sci = neg + 1 + (mxe + 1 if mxe > 0 else 0) + 2 + wexp
    fixed = w_fixed <= w_sci
    out = []
    for v in vals:
        if not np.isfinite(v):
            out.append("NaN" if np.isnan(v) else ("Inf" if v > 0 else "-Inf"))
        elif fixed:
            out.append(f"{v:.{rgt}f}")
        else:
            m, e = f"{v:.{mxe}e}".split("e")
            out.append(f"{m}e{e[0]}{int(e[1:]):0{wexp}d}")
    return out


def summary_text(tab: SummaryTable, n_samples: int, weighted: bool, names=None, digits: int = 4) -> str:
    """Text from the ``Data:`` line on (what ABC-DLS prints, ABC.py:844-861); the table as R's print.table shows it:
    digits = max(3, getOption("digits") - 3) = 4, one common format per column (format_r_column)."""
    names = names or tab.names or [f"P{p + 1}" for p in range(np.asarray(tab).shape[1])]
    lines = ["Data:", f" abc.out${'adj' if weighted else 'unadj'}.values ({n_samples} posterior samples)"]
    if weighted:
        lines += ["Weights:", " abc.out$weights"]
    lines.append("")
    a = np.asarray(tab)
    cols = [format_r_column(a[:, p], digits) for p in range(len(names))]
    widths = [max(len(names[p]), *(len(c) for c in cols[p])) for p in range(len(names))]
    lab = max(len(r) for r in ROWS)
    lines.append(" " * lab + " " + " ".join(n.rjust(w) for n, w in zip(names, widths)))
    for r in range(7):
        lines.append(ROWS[r].ljust(lab) + " " + " ".join(cols[p][r].rjust(widths[p]) for p in range(len(names))))
    return "\n".join(lines)
