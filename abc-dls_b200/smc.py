"""ABC-DLS SMC range rules — the P-sized host arithmetic between the table passes.

Same rules and order as ``ABC_DLS_SMC.updating_newrange`` (src/Classes/ABC.py:2921-2946),
``noise_injection_update`` (2881-2919) and ``lmrd_calculation`` (3016-3028), written on numpy arrays
instead of pandas frames.  The table-sized parts (posterior min/max, oldrange, box filter) are CUDA
kernels (csrc/smc.cu, csrc/loclinear.cu:k_col_stats).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class RangeTable:
    """One row per parameter: the reference's Newrange.csv columns (min, max, decrease)."""
    min: np.ndarray
    max: np.ndarray
    decrease: np.ndarray

    def to_csv(self, path: str, names) -> None:
        # reference: newrange.to_csv(folder + 'Newrange.csv', header=False)  (ABC.py:2729)
        with open(path, "w") as fh:
            for n, a, b, c in zip(names, self.min, self.max, self.decrease):
                fh.write(f"{n},{a!r},{b!r},{c!r}\n".replace("np.float64(", "").replace(")", ""))


def _ratio(new_min, new_max, old_min, old_max):
    with np.errstate(divide="ignore", invalid="ignore"):
        return (new_max - new_min) / (old_max - old_min)


def updating_newrange(post_min, post_max, old_min, old_max, decrease: float = 0.95, ndigits: int = 3) -> RangeTable:
    lo, hi = np.array(post_min, dtype=np.float64), np.array(post_max, dtype=np.float64)
    omin, omax = np.asarray(old_min, dtype=np.float64), np.asarray(old_max, dtype=np.float64)
    dec = _ratio(lo, hi, omin, omax)
    with np.errstate(invalid="ignore"):
        revert = ((dec > decrease) & (dec < 1)) | (dec == 0)
    lo[revert], hi[revert] = omin[revert], omax[revert]
    dec = _ratio(lo, hi, omin, omax)
    return RangeTable(np.round(lo, ndigits), np.round(hi, ndigits), np.round(dec, ndigits))


def noise_injection_update(rng: RangeTable, old_min, old_max, increase: float = 0.005,
                           hard_min: Optional[np.ndarray] = None, hard_max: Optional[np.ndarray] = None,
                           decrease: float = 0.95, ndigits: int = 3) -> RangeTable:
    lo, hi, dec = rng.min.astype(np.float64).copy(), rng.max.astype(np.float64).copy(), rng.decrease.copy()
    omin, omax = np.asarray(old_min, dtype=np.float64), np.asarray(old_max, dtype=np.float64)
    pad = (hi - lo) * increase * 0.5
    with np.errstate(invalid="ignore"):
        widen = dec > decrease
    lo = np.where(widen, lo - pad, lo)
    hi = np.where(widen, hi + pad, hi)
    if hard_min is not None and hard_max is not None:
        lo = np.maximum(np.asarray(hard_min, dtype=np.float64), lo)
        hi = np.minimum(np.asarray(hard_max, dtype=np.float64), hi)
    dec = _ratio(lo, hi, omin, omax)
    collapsed = np.round(hi, ndigits) == np.round(lo, ndigits)
    lo = np.where(collapsed, lo - 10.0 ** (-ndigits), lo)
    dec = np.where(collapsed, 1.0, dec)
    return RangeTable(np.round(lo, ndigits), np.round(hi, ndigits), np.round(dec, ndigits))


def lmrd_calculation(rng: RangeTable, hard_min, hard_max) -> float:
    new_w = np.abs(rng.max - rng.min)
    hard_w = np.abs(np.asarray(hard_max, dtype=np.float64) - np.asarray(hard_min, dtype=np.float64))
    return float(np.log(np.mean(new_w / hard_w)))
