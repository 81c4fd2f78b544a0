"""``summary.abc`` on the posterior sample (7 x P table: Min, 2.5%, Median, Mean, Mode, 97.5%, Max).

Min / Max / (weighted) Mean come from the ``k_col_stats`` kernel (they are what ABC-DLS's SMC reads,
ABC.py:2833, 2837).  The percentile rows are SURVEY §8f "next" work: they are computed here with a
device sort (R ``quantile`` type 7 for rejection; weighted order statistic = ``quantreg::rq(x ~ 1)`` for
loclinear).  The Mode row (argmax of R ``density``) is not implemented and reads NaN.
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


def summary_table(values: torch.Tensor, weights, stats: torch.Tensor, weighted: bool, intvl: float = 0.95, names=None):
    """values (n, P) global posterior sample, weights (n,) or None, stats (P, 6) from k_col_stats."""
    a = (1 - intvl) / 2
    n, P = values.shape
    out = torch.full((7, P), float("nan"), dtype=torch.float64, device=values.device)
    out[0], out[6] = stats[:, 0], stats[:, 1]
    if weighted:
        out[3] = stats[:, 3] / stats[:, 5]
        for p in range(P):
            xs, order = torch.sort(values[:, p], stable=True)
            cw = torch.cumsum(weights[order], 0)
            W = cw[-1]
            for row, tau in ((1, a), (2, 0.5), (5, 1 - a)):
                kk = torch.searchsorted(cw, (tau * W).reshape(1), right=False).clamp(max=n - 1)
                out[row, p] = xs[kk[0]]
    else:
        out[3] = stats[:, 2] / n
        for p in range(P):
            xs, _ = torch.sort(values[:, p])
            out[1, p], out[2, p], out[5, p] = _quantile7(xs, a), _quantile7(xs, 0.5), _quantile7(xs, 1 - a)
    return SummaryTable(out.cpu().numpy(), names)


def summary_text(tab: SummaryTable, n_samples: int, weighted: bool, names=None, digits: int = 4) -> str:
    """Text from the ``Data:`` line on (what ABC-DLS prints, ABC.py:844-861)."""
    names = names or tab.names or [f"P{p + 1}" for p in range(np.asarray(tab).shape[1])]
    lines = ["Data:", f" abc.out${'adj' if weighted else 'unadj'}.values ({n_samples} posterior samples)"]
    if weighted:
        lines += ["Weights:", " abc.out$weights"]
    lines.append("")
    cells = [[f"{v:.{digits}g}" for v in row] for row in np.asarray(tab)]
    widths = [max(len(names[p]), *(len(cells[r][p]) for r in range(7))) for p in range(len(names))]
    lab = max(len(r) for r in ROWS)
    lines.append(" " * lab + " " + " ".join(n.rjust(w) for n, w in zip(names, widths)))
    for r in range(7):
        lines.append(ROWS[r].ljust(lab) + " " + " ".join(c.rjust(w) for c, w in zip(cells[r], widths)))
    return "\n".join(lines)
