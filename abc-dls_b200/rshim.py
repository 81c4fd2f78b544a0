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
        return np.asarray(df).std(axis=0, ddof=1)   # R sd(): n-1 denominator (ABC.py:1778)


r = _R()
