"""Stand-in for the handful of ``rpy2.robjects.r[...]`` verbs ABC-DLS applies to abc results
(src/Classes/ABC.py:844-861, 1770-1780, 2821-2861): ``summary``, ``sink``, ``options``, ``pdf``,
``dev.off``, ``hist``, ``plot`` and ``sapply(df, 'sd')``.  Plot devices are no-ops (PDF output is out of
scope); ``summary`` prints the table text and returns it indexable like R's."""
from __future__ import annotations

import contextlib
import io
import sys

import numpy as np


class _R:
    def __init__(self):
        self._sink_file = None

    def options(self, **_kw):
        return None

    def __getitem__(self, verb: str):
        return getattr(self, "_" + verb.replace(".", "_"))

    def _summary(self, obj, **kw):
        # R's sink() diverts R's own output only: what summary() prints goes to the sink file (flushed at once - the
        # reference reads its first line while the sink is still open, ABC.py:1915-1918), Python's prints are untouched
        if self._sink_file is not None and not self._sink_file.closed:
            with contextlib.redirect_stdout(self._sink_file):
                res = obj.summary(**kw)
            self._sink_file.flush()
            return res
        return obj.summary(**kw)

    def _sink(self, path=None):  # R sink(file) / sink()
        if self._sink_file is not None:
            self._sink_file.close()
            self._sink_file = None
        if path is not None:
            self._sink_file = open(path, "w")

    def _pdf(self, *_a, **_k):
        return None

    def _dev_off(self, *_a, **_k):
        return None

    def _hist(self, *_a, **_k):
        return None

    def _plot(self, *_a, **_k):
        return None

    def _sapply(self, df, fn):
        if fn != "sd":
            raise NotImplementedError(fn)
        # R sd() (ABC.py:1778; ABC-DLS tests the result for == 0): n - 1 denominator around R's mean - a long double
        # sum with one refinement pass -, which is exactly c for a constant column; numpy's pairwise mean is not
        # (np.std(np.full(1000, 0.1), ddof=1) = 1.4e-17) and would let a constant column through
        a = np.asarray(df.to_numpy() if hasattr(df, "to_numpy") else df, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        n = a.shape[0]
        al = a.astype(np.longdouble)
        m = al.sum(axis=0) / np.longdouble(n)
        m = m + (al - m).sum(axis=0) / np.longdouble(n)
        dev = al - m
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sqrt(((dev * dev).sum(axis=0) / np.longdouble(n - 1))).astype(np.float64)


r = _R()
