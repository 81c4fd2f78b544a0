"""Pins the boundary and the SMC rules on THE REFERENCE ITSELF.

* CPU, only where ``/root/reference`` exists (this container): the reference's UNMODIFIED Python - ``abc_params``,
  ``abc_params_min_max`` + ``r_summary_min_max_*``, ``model_selection``, ``plot_power_of_ss``, ``goodness_fit``,
  ``plot_param_cv_error``, ``remove_constant`` (src/Classes/ABC.py) - runs over ``rabc`` + ``rshim`` bound the way
  INTEGRATION.md says, and what it prints / returns is checked against the oracle's numbers; its SMC range functions
  (ABC.py:2881-3028, src/SFS/Class.py:706-863) are the oracle for ``abc_dls_b200.smc``.
* anywhere (incl. the GPU box, where the reference does not exist): the committed fixtures that
  ``tests/golden/make_ref_golden.py`` generated from those runs.  ``-m gpu``: the CUDA engine replays the recorded
  abc() / postpr() / cv calls and must print the same text.
"""
import contextlib
import io
import json
import os
import re
import sys

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import refload  # noqa: E402
from oracle import abc_oracle as O  # noqa: E402
from abc_dls_b200 import rabc, smc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(not refload.available(), reason="/root/reference is not present on this box")
ROWS = ["Min.:", "2.5% Perc.:", "Median:", "Mean:", "Mode:", "97.5% Perc.:", "Max.:"]


# ---- helpers ---------------------------------------------------------------------------------------------------------
_NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def _ulp_of_print(tok: str) -> float:
    """Size of one unit in the last digit a printed number shows."""
    mant, _, exp = tok.lower().partition("e")
    dec = len(mant.split(".")[1]) if "." in mant else 0
    return 10.0 ** (-dec + (int(exp) if exp else 0))


def assert_same_text(got: str, want: str):
    """Identical token for token; numbers may differ by one unit of their last printed digit (two engines whose fp64
    results agree to 1e-11 can still round a 4-significant-digit print differently at a tie)."""
    gl = [l.split() for l in got.strip().splitlines()]
    wl = [l.split() for l in want.strip().splitlines()]
    assert len(gl) == len(wl), f"line count {len(gl)} != {len(wl)}\n{got}\n---\n{want}"
    for a, b in zip(gl, wl):
        assert len(a) == len(b), (a, b)
        for x, y in zip(a, b):
            if x == y:
                continue
            xs, ys = x.strip("[](),"), y.strip("[](),")
            assert _NUM.match(xs) and _NUM.match(ys), (a, b)
            assert abs(float(xs) - float(ys)) <= 1.01 * max(_ulp_of_print(xs), _ulp_of_print(ys)), (a, b)


def _capture(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = fn(*a, **k)
    return ret, buf.getvalue()


def _inputs():
    z = np.load(os.path.join(GOLD, "ref_callsites_inputs.npz"))
    g = json.load(open(os.path.join(GOLD, "ref_callsites.json")))
    names = g["names"]
    fr = dict(param=pd.DataFrame(z["par"], columns=names), pred=pd.DataFrame(z["pred"], columns=names),
              target=pd.DataFrame(z["target"], columns=names),
              index=pd.Series(np.array(g["models"])[z["mid"]]), ss_post=pd.DataFrame(z["ss_post"]),
              target_post=pd.DataFrame(z["target_post"]), mid=z["mid"])
    return z, g, fr


def _summary_blocks(text: str):
    """[(header tokens, 7 x ncol array of printed cells)] for every summary.abc table in ``text``."""
    lines = text.splitlines()
    out = []
    for i, l in enumerate(lines):
        if l.startswith("Min.:"):
            body = [lines[i + r] for r in range(7)]
            for lab, row in zip(ROWS, body):
                assert row.startswith(lab), (lab, row)
            cells = [row[len(lab):].split() for lab, row in zip(ROWS, body)]
            out.append((lines[i - 1].split(), cells))
    return out


@pytest.fixture()
def cpu_boundary():
    """rabc bound to an engine whose stages are the numpy double (no GPU here), seeded cv draws."""
    from abc_dls_b200.engine import AbcEngine
    from cpu_stages import CpuStages
    rabc.set_engine(AbcEngine(stages=CpuStages()))
    rabc.default_seed = 4242
    yield
    rabc.set_engine(None)
    rabc.default_seed = None


@pytest.fixture()
def ref(cpu_boundary, tmp_path, monkeypatch):
    ABC, SC = refload.load()
    monkeypatch.chdir(tmp_path)            # the reference drops its sink files into the working directory
    ABC.robjects.r["sink"]()
    yield ABC, SC
    ABC.robjects.r["sink"]()
    ABC.abc = rabc


# ---- A. the reference's own call sites over rabc + rshim (CPU, this container) ------------------------------------------
@needs_ref
@pytest.mark.parametrize("method", ["rejection", "loclinear"])
def test_unmodified_abc_params_prints_the_oracle_summary(ref, method):
    """ABC.py:1929-1973: P separate abc.abc(DataFrame, DataFrame, DataFrame, method=, tol=) calls + r_summary (sink ->
    summary -> sink() -> print from 'Data:'), then for rejection one joint call."""
    ABC, _ = ref
    z, g, fr = _inputs()
    _, text = _capture(ABC.ABC_DLS_Params.abc_params, target=fr["target"], param=fr["param"], ss=fr["pred"],
                       tol=g["tol"], method=method, name="post.pdf")
    assert_same_text(text, g[f"abc_params_{method}"])         # the committed fixture is what the reference prints today
    blocks = _summary_blocks(text)
    par, pred, tgt = z["par"], z["pred"].astype(np.float64), z["target"][0]
    P = par.shape[1]
    assert len(blocks) == P + (1 if method == "rejection" else 0)
    for c in range(P):                                          # per-column problems: d = P = 1
        o = O.abc(tgt[c:c + 1], par[:, c], pred[:, c:c + 1], g["tol"], method)
        want = O.r_format_real(O.summary_abc(o)[:, 0])
        head, cells = blocks[c]
        assert head == [g["names"][c]] and [r[0] for r in cells] == want
        assert f"({o.nacc} posterior samples)" in text
    if method == "rejection":                                   # "together": one joint N x P problem
        o = O.abc(tgt, par, pred, g["tol"], method)
        tab = O.summary_abc(o)
        head, cells = blocks[P]
        assert head == g["names"]
        for c in range(P):
            assert [r[c] for r in cells] == O.r_format_real(tab[:, c])
    assert ("Weights:" in text) == (method == "loclinear")


@needs_ref
@pytest.mark.parametrize("method", ["rejection", "loclinear"])
def test_unmodified_abc_params_min_max_reads_summary_0_and_6(ref, method):
    """ABC.py:2788-2840: summary(res)[0] / summary(res)[6] per column -> the SMC posterior box."""
    ABC, _ = ref
    z, g, fr = _inputs()
    (mn, mx), printed = _capture(ABC.ABC_DLS_SMC.abc_params_min_max, target=fr["target"], param=fr["param"],
                                 ss=fr["pred"], tol=g["tol"], method=method)
    assert printed == ""                                        # both summary() calls were sunk to the temp file
    par, pred, tgt = z["par"], z["pred"].astype(np.float64), z["target"][0]
    for c in range(par.shape[1]):
        o = O.abc(tgt[c:c + 1], par[:, c], pred[:, c:c + 1], g["tol"], method)
        v = o.unadj_values if method == "rejection" else o.adj_values
        np.testing.assert_allclose([mn[c], mx[c]], [v.min(), v.max()], rtol=0 if method == "rejection" else 1e-9)
    np.testing.assert_allclose([list(mn), list(mx)], g[f"minmax_{method}"], rtol=1e-12)


@needs_ref
def test_unmodified_r_summary_min_max_all_reads_first_and_last_row(ref):
    """ABC.py:2842-2861 (the joint branch; ABC-DLS reaches it with method='neuralnet', the table logic is the same)."""
    ABC, _ = ref
    z, g, fr = _inputs()
    res = rabc.abc(target=fr["target"], param=fr["param"], sumstat=fr["pred"], method="loclinear", tol=g["tol"])
    (mn, mx), printed = _capture(ABC.ABC_DLS_SMC.r_summary_min_max_all, res)
    o = O.abc(z["target"][0], z["par"], z["pred"].astype(np.float64), g["tol"], "loclinear")
    np.testing.assert_allclose(mn.to_numpy(), o.adj_values.min(0), rtol=1e-9)
    np.testing.assert_allclose(mx.to_numpy(), o.adj_values.max(0), rtol=1e-9)
    assert printed == ""


@needs_ref
def test_unmodified_model_selection_prints_the_oracle_proportions(ref):
    """ABC.py:901-921: abc.postpr(target=DataFrame, index=Series[str], sumstat=DataFrame, tol=, method=) + r_summary."""
    ABC, _ = ref
    z, g, fr = _inputs()
    _, text = _capture(ABC.ABC_DLS_Classification.model_selection, target=fr["target_post"], ss=fr["ss_post"],
                       index=fr["index"], tol=g["tol"], method="rejection")
    assert_same_text(text, g["model_selection_rejection"])
    o = O.postpr(z["target_post"][0], np.array(g["models"])[z["mid"]], z["ss_post"], g["tol"])
    lines = text.splitlines()
    k = lines.index("Proportion of accepted simulations (rejection):")
    assert lines[k + 1].split() == list(o.models) and lines[k + 2].split() == O.r_format_real(o.pred)
    b = lines.index("Bayes factors:")
    bf = o.bayes_factors()
    for i, m in enumerate(o.models):
        assert lines[b + 2 + i].split() == [m] + [O.r_format_real(bf[:, j])[i] for j in range(len(o.models))]
    assert f"({int(o.counts.sum())} posterior samples)" in text


@needs_ref
def test_unmodified_plot_power_of_ss_and_goodness_fit(ref):
    """ABC.py:863-899 (abc.cv4postpr(index=, sumstat=, nval=, tol=, method=), first sunk line + the summary object) and
    946-965 (abc.gfit(target, ss, repeats, tol=) positional)."""
    import make_ref_golden as M
    ABC, _ = ref
    z, g, fr = _inputs()
    rec = M.Recorder(rabc)
    ABC.abc = rec
    _, text = _capture(ABC.ABC_DLS_Classification.plot_power_of_ss, ss=fr["ss_post"], index=fr["index"], tol=g["tol"],
                       repeats=4, method="rejection")
    assert_same_text(text, g["plot_power_of_ss_rejection"])
    cv = rec.log[-1][1]
    assert cv.cvsamples.tolist() == g["cv4postpr_cvsamples"]
    labels = np.array(g["models"])[z["mid"]]
    o = O.cv4postpr(labels, z["ss_post"], 4, g["tol"], cvsamples=cv.cvsamples)
    s = O.summary_cv4postpr(o)
    np.testing.assert_array_equal(cv.summary(print_=False)["conf.matrix"][g["tol"]], s["conf.matrix"][g["tol"]])
    np.testing.assert_allclose(cv.summary(print_=False)["probs"][g["tol"]], s["probs"][g["tol"]], rtol=1e-12)
    assert text.splitlines()[0] == "Confusion matrix based on 4 samples for each model."

    sub = fr["ss_post"].iloc[z["mid"] == 1]
    _, text = _capture(ABC.ABC_DLS_Classification.goodness_fit, target=fr["target_post"], ss=sub, name="MNDX",
                       tol=g["tol"], repeats=5)
    assert_same_text(text, g["goodness_fit"])
    fit = rec.log[-1][1]
    # the held-out rows are drawn by cv._draw(N, 5, seed): hand the same rows to the oracle
    from abc_dls_b200 import cv as _cv
    og = O.gfit(z["target_post"][0], sub.to_numpy(), 5, g["tol"], samples=_cv._draw(len(sub), 5, rabc.default_seed))
    np.testing.assert_allclose(fit.dist_sim, og.dist_sim, rtol=1e-12)
    assert abs(fit.dist_obs - og.dist_obs) <= 1e-12 * og.dist_obs
    assert text.splitlines()[0] == "MNDX" and f"[1] {O.summary_gfit(og)['pvalue']:.4g}" in text


@needs_ref
@pytest.mark.parametrize("method", ["rejection", "loclinear"])
def test_unmodified_plot_param_cv_error(ref, method):
    """ABC.py:1877-1927: per-column abc.cv4abc(param=DataFrame, sumstat=DataFrame, nval=, tols=, method=, trace=) +
    summary(cv)[0][0]; for rejection also the joint call, whose summary is read while the sink is still open."""
    import make_ref_golden as M
    ABC, _ = ref
    z, g, fr = _inputs()
    rec = M.Recorder(rabc)
    ABC.abc = rec
    _, text = _capture(ABC.ABC_DLS_Params.plot_param_cv_error, param=fr["param"], ss=fr["pred"], tol=g["tol"],
                       repeats=5, method=method, name="cv.pdf")
    assert_same_text(text, g[f"plot_param_cv_error_{method}"])
    cvs = [r for n, r in rec.log if n == "cv4abc"]
    par, pred = z["par"], z["pred"].astype(np.float64)
    P = par.shape[1]
    assert len(cvs) == P + (1 if method == "rejection" else 0)
    printed = [float(l) for l in text.splitlines() if _NUM.match(l.strip()) and "." in l and len(l) > 12]
    for c in range(P):
        o = O.cv4abc(par[:, c], pred[:, c], 5, g["tol"], method, cvsamples=cvs[c].cvsamples)
        want = O.summary_cv4abc(o)[0, 0]
        assert abs(cvs[c].prediction_error()[0, 0] - want) <= 1e-9 * want
        assert abs(printed[c] - want) <= 1e-9 * want           # print(summary[0][0])
    if method == "rejection":
        o = O.cv4abc(par, pred, 5, g["tol"], method, cvsamples=cvs[P].cvsamples)
        np.testing.assert_allclose(cvs[P].prediction_error(), O.summary_cv4abc(o), rtol=1e-9)


@needs_ref
def test_unmodified_remove_constant_drops_exactly_the_zero_sd_columns(ref):
    """ABC.py:1770-1823 over rshim's sapply(df, 'sd'): R's sd is exactly 0 for a constant column (numpy's is 1e-17 for
    0.1) and non-zero for one that varies in its last bit."""
    ABC, _ = ref
    z, g, fr = _inputs()
    pred = fr["pred"].astype(np.float64).copy()
    pred["T_split"] = 0.1
    pred["mig"] = 0.1
    pred.loc[17, "mig"] = np.nextafter(0.1, 1.0)
    (tp, p4, pu), text = _capture(ABC.ABC_DLS_Params.remove_constant, test_predictions=pred,
                                  predict4mreal=fr["target"].copy(), params_unscaled=fr["param"].copy())
    assert list(tp.columns) == ["N_A", "mig"] and list(p4.columns) == ["N_A", "mig"]
    assert list(pu.columns) == g["names"]          # the reference discards its own drop of the parameter column (A11)
    assert "T_split" in text


# ---- B. SMC rules: the reference's functions are the oracle of abc_dls_b200.smc -------------------------------------------
def _rules():
    return np.load(os.path.join(GOLD, "ref_smc_rules.npz"))


def test_smc_rules_reproduce_the_reference_fixture_bit_for_bit():
    z = _rules()
    for i in range(int(z["n_cases"][0])):
        g = lambda k: z[f"c{i}_{k}"]
        dec, inc, hard = g("cfg")
        r = smc.updating_newrange(g("pmin"), g("pmax"), g("omin"), g("omax"), dec)
        np.testing.assert_array_equal(np.stack([r.min, r.max, r.decrease], 1), g("upd"))
        o = O.updating_newrange(g("pmin"), g("pmax"), g("omin"), g("omax"), dec)
        np.testing.assert_array_equal(np.stack(o, 1), g("upd"))
        if inc > 0:
            hm = (g("hmin"), g("hmax")) if hard else (None, None)
            r = smc.noise_injection_update(r, g("omin"), g("omax"), inc, hm[0], hm[1], dec)
            np.testing.assert_array_equal(np.stack([r.min, r.max, r.decrease], 1), g("inj"))
            o = O.noise_injection_update(*o, g("omin"), g("omax"), inc, hm[0], hm[1], dec)
            np.testing.assert_array_equal(np.stack(o, 1), g("inj"))
        assert smc.lmrd_calculation(r, g("hmin"), g("hmax")) == g("lmrd")[0]
    np.testing.assert_array_equal(O.narrowing_params(z["narrow_par"], z["narrow_lo"], z["narrow_hi"]), z["narrow_lines"])


def test_snakemake_glue_reproduces_the_reference_fixture(tmp_path):
    """src/SFS/Class.py:706-796, 831-863: median of the repeats' Newrange.csv, update with a name-aligned hard range,
    the csv it writes, lmrd4mcsv, remove_repeated_params."""
    z = _rules()
    names = [str(n) for n in z["snk_names"]]
    files = []
    for r, tab in enumerate(z["snk_files"]):
        f = tmp_path / f"Newrange_{r}.csv"
        f.write_text("".join(f"{n},{a!r},{b!r},0.5\n" for n, (a, b) in zip(names, tab.tolist())))
        files.append(str(f))
    got_names, lo, hi = smc.median_newrange(files)
    assert got_names == names
    np.testing.assert_array_equal(np.stack([lo, hi], 1), z["snk_median"])
    case, dec, inc = z["snk_case"]
    i = int(case)
    order = z["snk_hard_order"]
    hf = tmp_path / "hardrange.csv"
    hf.write_text("".join(f"{names[j]},{float(z[f'c{i}_hmin'][j])!r},{float(z[f'c{i}_hmax'][j])!r}\n" for j in order))
    out = tmp_path / "Newrange.csv"
    _, rng = smc.multiple_newrange2updatednewrange(files, z[f"c{i}_omin"], z[f"c{i}_omax"], dec, inc, str(hf), str(out))
    np.testing.assert_array_equal(np.stack([rng.min, rng.max, rng.decrease], 1), z["snk_updated"])
    assert out.read_text() == str(z["snk_newrange_csv"])        # byte-identical to DataFrame.to_csv(header=False)
    assert smc.lmrd4mcsv(str(hf), str(out)) == z["snk_lmrd"][0]
    allcsv = tmp_path / "All.csv"
    allcsv.write_text(str(z["dedup_in"]))
    smc.remove_repeated_params(str(allcsv), 2, str(tmp_path / "dedup.csv"))
    assert (tmp_path / "dedup.csv").read_text() == str(z["dedup_out"])
    smc.remove_repeated_params(str(allcsv), 2, str(tmp_path / "first.csv"), as_reference=False)
    rows = [l.split(",") for l in (tmp_path / "first.csv").read_text().splitlines()]
    assert len({(a, b) for a, b, _ in rows}) == len(rows) == len({tuple(l.split(",")[:2])
                                                                  for l in str(z["dedup_in"]).splitlines()})


@needs_ref
def test_fixtures_are_what_the_reference_produces_today(tmp_path):
    """Re-runs the generator's SMC part against the live reference and compares with the committed file."""
    import make_ref_golden as M
    ABC, SC = refload.load()
    M.make_smc_rules(ABC, SC, str(tmp_path / "rules.npz"))
    a, b = np.load(tmp_path / "rules.npz"), _rules()
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)


@needs_ref
def test_narrowed_csv_against_the_reference(ref, tmp_path):
    """src/SFS/Class.py:799-829 (header-less All.csv) and ABC.py:2948-3014 (csv with a header): same SET of lines as the
    reference writes (both shuffle).  Integer parameter names: with pandas >= 3 the reference's positional
    ``parmin[index]`` on a name-indexed Series no longer works, so the range file labels the rows 0..P-1."""
    ABC, SC = ref
    rng = np.random.default_rng(3)
    n, P = 400, 3
    par = np.round(rng.uniform(0, 1, (n, P)), 2)
    lines = [",".join(repr(float(v)) for v in par[i]) + f",{i},{i * 7 % 13}" for i in range(n)]
    (tmp_path / "All.csv").write_text("\n".join(lines) + "\n")
    lo, hi = np.array([0.1, 0.2, 0.0]), np.array([0.8, 0.9, 0.55])
    (tmp_path / "Newrange.csv").write_text("".join(f"{j},{float(lo[j])!r},{float(hi[j])!r},0.5\n" for j in range(P)))
    (tmp_path / "ref").mkdir()
    # the reference shells out to its line shuffler with a bash process substitution: run it, fall back to the
    # unshuffled Narrows.csv it would have shuffled if the shell tool is unavailable
    try:
        n_ref = SC.ABC_DLS_SMC_Snakemake.narrowing_input(paramsnumbers=P, inputfile=str(tmp_path / "All.csv"),
                                                        rangefile=str(tmp_path / "Newrange.csv"),
                                                        folder=str(tmp_path / "ref") + "/")
        want = sorted((tmp_path / "ref" / "Narrowed.csv").read_text().splitlines())
    except SystemExit:
        want = sorted((tmp_path / "ref" / "Narrows.csv").read_text().splitlines())
        n_ref = len(want)
    (tmp_path / "mine").mkdir()
    n_mine = smc.narrowing_input_rangefile(P, str(tmp_path / "All.csv"), str(tmp_path / "Newrange.csv"),
                                           folder=str(tmp_path / "mine") + "/", engine=rabc.engine(), seed=1)
    got = (tmp_path / "mine" / "Narrowed.csv").read_text().splitlines()
    keep = np.all((par >= lo) & (par <= hi), axis=1)
    assert n_mine == n_ref == int(keep.sum()) and sorted(got) == want
    assert got != sorted(got, key=lambda l: int(l.split(",")[3]))          # shuffled
    # header variant (ABC.py:2974-2990: line numbers are index + 2)
    (tmp_path / "sims.csv").write_text("a,b,c,id,x\n" + "\n".join(lines) + "\n")
    ln_ref = ABC.ABC_DLS_SMC.narrowing_params(params=pd.DataFrame(par), parmin=pd.Series(lo), parmax=pd.Series(hi))
    ABC.ABC_DLS_SMC.extracting_by_linenumber(file=str(tmp_path / "sims.csv"), linenumbers=ln_ref,
                                             outputfile=str(tmp_path / "ref_rows.csv"))
    smc.narrowing_input(str(tmp_path / "sims.csv"), lo, hi, P, folder=str(tmp_path / "mine") + "/h_", header=True,
                        engine=rabc.engine(), shuffle=False)
    assert (tmp_path / "mine" / "h_Narrowed.csv").read_text() == (tmp_path / "ref_rows.csv").read_text()


# ---- C. the CUDA engine replays the recorded calls (GPU box: no reference there) ----------------------------------------
@pytest.fixture()
def cuda_boundary(engine):
    rabc.set_engine(engine)
    rabc.default_seed = 4242
    yield engine
    rabc.set_engine(None)
    rabc.default_seed = None


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["rejection", "loclinear"])
def test_cuda_engine_prints_what_the_reference_call_sites_printed(cuda_boundary, method):
    """Same calls, same argument types as ABC.py:1949-1953 / 1964 / 2806-2809, on the CUDA engine; the text blocks
    must equal the ones the unmodified reference printed over the shim (fixture), the min / max its SMC read."""
    z, g, fr = _inputs()
    want_blocks = _summary_blocks(g[f"abc_params_{method}"])
    P = len(g["names"])
    mins, maxs = [], []
    for c in range(P):
        res = rabc.abc(target=pd.DataFrame(fr["target"].iloc[:, c]), param=pd.DataFrame(fr["param"].iloc[:, c]),
                       sumstat=pd.DataFrame(fr["pred"].iloc[:, c]), method=method, tol=g["tol"])
        tab, text = _capture(res.summary)
        head, cells = _summary_blocks(text)[0]
        assert head == want_blocks[c][0]
        assert_same_text("\n".join(" ".join(r) for r in cells), "\n".join(" ".join(r) for r in want_blocks[c][1]))
        mins.append(tab[0])
        maxs.append(tab[6])
    np.testing.assert_allclose([mins, maxs], g[f"minmax_{method}"], rtol=1e-9)
    if method == "rejection":
        res = rabc.abc(target=fr["target"], param=fr["param"], sumstat=fr["pred"], method=method, tol=g["tol"])
        _, text = _capture(res.summary)
        head, cells = _summary_blocks(text)[0]
        assert head == want_blocks[P][0]
        assert_same_text("\n".join(" ".join(r) for r in cells), "\n".join(" ".join(r) for r in want_blocks[P][1]))


@pytest.mark.gpu
def test_cuda_engine_postpr_cv_gfit_print_what_the_reference_call_sites_printed(cuda_boundary):
    z, g, fr = _inputs()
    fit = rabc.postpr(target=fr["target_post"], index=fr["index"], sumstat=fr["ss_post"], tol=g["tol"],
                      method="rejection")
    assert_same_text(fit.summary_text(), g["model_selection_rejection"])
    cv = rabc.cv4postpr(index=fr["index"], sumstat=fr["ss_post"], nval=4, tol=g["tol"], method="rejection")
    assert cv.cvsamples.tolist() == g["cv4postpr_cvsamples"]
    _, text = _capture(cv.summary)
    assert text.splitlines()[0] == g["plot_power_of_ss_rejection"].splitlines()[0]
    assert_same_text(repr(cv.summary(print_=False)), "\n".join(g["plot_power_of_ss_rejection"].splitlines()[1:]))
    gf = rabc.gfit(fr["target_post"], fr["ss_post"].iloc[z["mid"] == 1], 5, tol=g["tol"])
    np.testing.assert_allclose(gf.dist_sim, g["gfit_dist_sim"], rtol=1e-12)
    for method in ("rejection", "loclinear"):
        want = [float(l) for l in g[f"plot_param_cv_error_{method}"].splitlines()
                if _NUM.match(l.strip()) and "." in l and len(l) > 12]
        for c in range(3):
            r = rabc.cv4abc(param=pd.DataFrame(fr["param"].iloc[:, c]), sumstat=pd.DataFrame(fr["pred"].iloc[:, c]),
                            nval=5, tols=g["tol"], method=method, trace=False)
            assert r.cvsamples.tolist() == g[f"cv4abc_cvsamples_{method}"][c]
            assert abs(r.summary(print_=False)[0][0] - want[c]) <= 1e-9 * want[c]
