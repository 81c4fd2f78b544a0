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


# ---- cross-round file formats of the Snakemake loop ------------------------------------------------
# Newrange.csv / hardrange.csv: no header, one row per parameter: name, min, max[, decrease]  (ABC.py:2729;
# src/SFS/Class.py:706-796, 850-863).  P-sized host work; the table-sized parts stay with engine.oldrange / box_filter.
def read_range_csv(path: str):
    """-> (names, min, max) of a header-less range file (Newrange.csv, hardrange.csv)."""
    names, lo, hi = [], [], []
    with open(path) as fh:
        for line in fh:
            cells = line.rstrip("\n").split(",")
            if len(cells) < 3 or not cells[0]:
                continue
            names.append(cells[0])
            lo.append(float(cells[1]))
            hi.append(float(cells[2]))
    return names, np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)


def median_newrange(files):
    """Per-parameter median of the min and of the max over the Newrange.csv files of ``abc_repeats`` independent
    repeats (src/SFS/Class.py:736-749; an even number of files averages the two middle values, like pandas)."""
    tabs = [read_range_csv(f) for f in files]
    names = tabs[0][0]
    return names, np.median(np.stack([t[1] for t in tabs]), axis=0), np.median(np.stack([t[2] for t in tabs]), axis=0)


def multiple_newrange2updatednewrange(newrange_files, old_min, old_max, decrease: float = 0.95, increase: float = 0.0,
                                      hardrange_file: Optional[str] = None, outfile: Optional[str] = "Newrange.csv"):
    """src/SFS/Class.py:706-734: median over the repeats -> updating_newrange -> [noise_injection_update] -> csv.
    ``old_min`` / ``old_max`` are the per-parameter extremes of the simulated table (``engine.oldrange`` on the device
    table replaces the reference's second read of the csv, Class.py:751-763)."""
    names, lo, hi = median_newrange(newrange_files)
    rng = updating_newrange(lo, hi, old_min, old_max, decrease)
    if increase > 0:
        hmin = hmax = None
        if hardrange_file:
            hnames, hmin, hmax = read_range_csv(hardrange_file)
            order = [hnames.index(n) for n in names]            # pandas aligns the hard range on the parameter names
            hmin, hmax = hmin[order], hmax[order]
        rng = noise_injection_update(rng, old_min, old_max, increase, hmin, hmax, decrease)
    if outfile:
        rng.to_csv(outfile, names)
    return names, rng


def lmrd4mcsv(hardrange_file: str, newrange_file: str) -> float:
    """src/SFS/Class.py:850-863: log of the mean range decrease straight from the two files."""
    hnames, hmin, hmax = read_range_csv(hardrange_file)
    names, lo, hi = read_range_csv(newrange_file)
    order = [hnames.index(n) for n in names]
    return lmrd_calculation(RangeTable(lo, hi, np.zeros_like(lo)), hmin[order], hmax[order])


def csv_linenumbers(rows, header: bool = True) -> np.ndarray:
    """0-based table rows (``engine.box_filter``) -> the 1-based csv line numbers the reference extracts:
    row + 2 for a file with a header line (ABC.py:2974-2990), row + 1 without (src/SFS/Class.py:817-818)."""
    rows = np.asarray(rows.cpu() if hasattr(rows, "cpu") else rows, dtype=np.int64)
    return rows + (2 if header else 1)


# ---- Narrowed.csv: the simulations that stay inside the new box are reused by the next round ------------------------
# ABC.py:2948-3014 (narrowing_input / narrowing_params / extracting_by_linenumber, csv WITH a header line) and the
# Snakemake variant src/SFS/Class.py:799-846 (All.csv WITHOUT a header; remove_repeated_params).  The row filter is the
# device box filter (csrc/smc.cu:k_box_tiles through engine.box_filter); reading and writing text lines is host I/O.
def _open_text(path: str):
    if path.endswith(".gz"):
        import gzip
        return gzip.open(path, "rt")
    return open(path, "r")


def extracting_by_linenumber(file: str, linenumbers, outputfile: str = "out.txt") -> str:
    """ABC.py:2992-3014: copy the lines whose 1-BASED number is in ``linenumbers`` (any order, duplicates ignored),
    in file order.  A ``.gz`` input is read through gzip instead of the reference's ``zcat > temp.csv``."""
    want = np.unique(np.asarray(linenumbers, dtype=np.int64).reshape(-1)) - 1
    want = want[want >= 0]
    k = 0
    with _open_text(file) as src, open(outputfile, "w") as out:
        for lno, ln in enumerate(src):
            if k >= want.size:
                break
            if lno == want[k]:
                out.write(ln)
                k += 1
    return outputfile


def _read_params(inputfile: str, paramsnumbers: int, header: bool) -> np.ndarray:
    import pandas
    tab = pandas.read_csv(inputfile, usecols=range(paramsnumbers), header=0 if header else None)
    return np.ascontiguousarray(tab.to_numpy(dtype=np.float64))


def _shuffle_lines(src: str, dst: str, seed) -> int:
    """The reference pipes the extracted lines through its external-memory shuffler (ABC.py:331-364, shuffle.py); the
    order is random there too, so only the SET of lines is comparable.  ``seed`` makes it reproducible."""
    with open(src) as fh:
        lines = fh.readlines()
    order = np.random.default_rng(seed).permutation(len(lines))
    with open(dst, "w") as fh:
        fh.writelines(lines[i] for i in order)
    return len(lines)


def narrowing_input(inputfile: str, new_min, new_max, paramsnumbers: int, folder: str = "", header: bool = True,
                    engine=None, seed=None, shuffle: bool = True) -> int:
    """Writes ``folder + 'Narrowed.csv'``: the lines of ``inputfile`` whose first ``paramsnumbers`` columns all lie in
    [new_min, new_max] (inclusive, like ``Series.between``), shuffled; returns its line count.
    ``header=True``: ABC.py:2948-2972 (the csv named in the info file; its header line is never copied);
    ``header=False``: src/SFS/Class.py:799-829 (All.csv).  The filter runs on the device (``engine.box_filter``)."""
    import os
    import torch
    out = folder + "Narrowed.csv"
    if os.path.getsize(inputfile) == 0:
        open(out, "w").close()
        return 0
    if engine is None:
        from . import rabc
        engine = rabc.engine()
    par = _read_params(inputfile, paramsnumbers, header)
    tab = torch.from_numpy(par.T.copy()).to(engine.device)            # (P, n) column-major table
    rows = engine.box_filter(tab, torch.as_tensor(np.asarray(new_min, dtype=np.float64)),
                             torch.as_tensor(np.asarray(new_max, dtype=np.float64)))
    tmp = folder + "Narrows.csv"
    extracting_by_linenumber(inputfile, csv_linenumbers(rows, header=header), tmp)
    if shuffle:
        n = _shuffle_lines(tmp, out, seed)
        os.remove(tmp)
    else:
        os.replace(tmp, out)
        with open(out) as fh:
            n = sum(1 for _ in fh)
    return n


def narrowing_input_rangefile(paramsnumbers: int, inputfile: str, rangefile: str, folder: str = "", engine=None,
                              seed=None) -> int:
    """src/SFS/Class.py:799-829 with its own argument order: the new box comes from a Newrange.csv."""
    _, lo, hi = read_range_csv(rangefile)
    return narrowing_input(inputfile, lo, hi, paramsnumbers, folder, header=False, engine=engine, seed=seed)


def remove_repeated_params(inputfile: str, paramsnumbers: int, outputfile: str, as_reference: bool = True) -> str:
    """src/SFS/Class.py:831-846: one line per distinct parameter vector of a header-less All.csv.

    The reference takes the 0-based rows ``params.drop_duplicates().index`` (first occurrence of every vector) and
    hands them to ``extracting_by_linenumber``, which treats its argument as 1-based line numbers: what it actually
    writes is the line BEFORE every first occurrence (row r - 1; nothing for r = 0).  ``as_reference=True`` (default)
    reproduces exactly that file, so a pipeline that swaps engines sees identical bytes; ``as_reference=False`` writes
    the first occurrences themselves (what the docstring there describes)."""
    par = _read_params(inputfile, paramsnumbers, header=False)
    _, first = np.unique(par, axis=0, return_index=True)
    first = np.sort(first)                                        # drop_duplicates keeps file order
    return extracting_by_linenumber(inputfile, first if as_reference else first + 1, outputfile)
