"""cv4abc / cv4postpr / gfit: the leave-one-out sequencing against the oracle with the same held-out
rows.  CPU tests run the host logic (``cv.py``: in-place hole buffers, point estimates, summaries) over the numpy stage
double; the ``gpu`` tests run the same comparisons through the CUDA engine and the rpy2-shaped facade."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import abc_oracle as O
from abc_dls_b200 import cv as CV
from abc_dls_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _cpu_engine():
    from abc_dls_b200.engine import AbcEngine
    from cpu_stages import CpuStages
    return AbcEngine(stages=CpuStages())


def test_leave_one_out_buffers_follow_any_hole_sequence():
    rng = np.random.default_rng(3)
    for n in (3, 8, 101):
        a = torch.from_numpy(rng.normal(size=(3, n)))
        b = torch.from_numpy(rng.integers(0, 5, size=(1, n)).astype(np.int32))
        loo = CV.LeaveOneOut([a, b])
        for s in [0, n - 1, 0, 1, n - 2] + rng.integers(0, n, 20).tolist():
            ba, bb = loo.hold_out(int(s))
            np.testing.assert_array_equal(ba.numpy(), np.delete(a.numpy(), s, axis=1))
            np.testing.assert_array_equal(bb.numpy(), np.delete(b.numpy(), s, axis=1))
            assert ba.stride(1) == 1 and ba.shape == (3, n - 1)
    with pytest.raises(CV.AbcError):
        CV.LeaveOneOut([torch.zeros(1, 2, dtype=torch.float64)])


def _check_cv4abc(eng, method, statistic, N=400, P=2, nval=6, tols=(0.2, 0.1)):
    par, ss, _ = synth.abc_tables(N, P, P, "cpu")
    samp = np.random.default_rng(11).choice(N, nval, replace=False)
    dev = eng.device
    got = CV.cv4abc(eng, par.to(dev), ss.to(dev), nval, tols, method, statistic=statistic, cvsamples=samp)
    want = O.cv4abc(par.numpy().T, ss.numpy().T, nval, tols, method, statistic=statistic, cvsamples=samp)
    np.testing.assert_array_equal(got.tols, want.tols)
    np.testing.assert_array_equal(got.true, want.true)
    scale = np.abs(want.true).max(axis=0)
    for t in want.tols:
        assert np.max(np.abs(got.estim[float(t)] - want.estim[float(t)]) / scale) < 1e-9
    np.testing.assert_allclose(got.summary(print_=False), O.summary_cv4abc(want), rtol=1e-7)
    assert got.summary_text().startswith("Prediction error based on a cross-validation sample of %d" % nval)
    return got


@pytest.mark.parametrize("method,statistic", [("rejection", "median"), ("loclinear", "median"), ("loclinear", "mean"),
                                              ("rejection", "mode")])
def test_cv4abc_host_logic_matches_oracle(method, statistic):
    _check_cv4abc(_cpu_engine(), method, statistic)


def test_cv4postpr_and_gfit_host_logic_match_oracle():
    eng = _cpu_engine()
    N, nval = 600, 3
    ids, ss, _ = synth.postpr_tables(N, "cpu")
    labels = np.array(["BNDX", "MNDX", "SNDX"])[ids.numpy()]
    rng = np.random.default_rng(5)
    samp = np.concatenate([rng.choice(np.flatnonzero(ids.numpy() == k), nval, replace=False) for k in range(3)])
    got = CV.cv4postpr(eng, ids, ["BNDX", "MNDX", "SNDX"], ss, nval, [0.1, 0.05], cvsamples=samp)
    want = O.cv4postpr(labels, ss.numpy().T, nval, [0.1, 0.05], cvsamples=samp)
    assert got.true == want.true.tolist()
    for t in want.tols:
        np.testing.assert_array_equal(got.estim[float(t)], want.estim[float(t)])          # counts / nacc: exact
    gs, ws = got.summary(print_=False), O.summary_cv4postpr(want)
    for t in want.tols:
        np.testing.assert_array_equal(gs["conf.matrix"][float(t)], ws["conf.matrix"][float(t)])
        np.testing.assert_allclose(gs["probs"][float(t)], ws["probs"][float(t)], rtol=1e-12)
    assert got.summary_text().startswith("Confusion matrix based on 3 samples for each model.")

    par, ss2, tgt = synth.abc_tables(500, 3, 3, "cpu")
    s2 = rng.choice(500, 8, replace=False)
    for statistic in ("mean", "median"):
        g = CV.gfit(eng, tgt, ss2, 8, 0.1, statistic=statistic, samples=s2)
        w = O.gfit(tgt.numpy(), ss2.numpy().T, 8, 0.1, statistic=statistic, samples=s2)
        np.testing.assert_allclose(g.dist_sim, w.dist_sim, rtol=1e-12)
        assert abs(g.dist_obs - w.dist_obs) <= 1e-12 * w.dist_obs
        a, b = g.summary(print_=False), O.summary_gfit(w)
        assert a["pvalue"] == b["pvalue"]
        np.testing.assert_allclose(a["s.dist.sim"], b["s.dist.sim"], rtol=1e-12)


def test_rshim_verbs(tmp_path):
    """robjects.r[...] verbs ABC-DLS uses around the results (ABC.py:887-896, 1778, 1896-1925)."""
    from abc_dls_b200.rshim import r

    class Obj:
        def summary(self, **kw):
            print("first line")
            return np.array([[0.25]])

    r.options(width=10000)
    p = tmp_path / "sink.txt"
    r["sink"](str(p))
    x = r["summary"](Obj())
    assert open(p).readline() == "first line\n"          # readable while the sink is still open (ABC.py:1915-1918)
    print("python output is not diverted")
    r["sink"]()
    assert open(p).read() == "first line\n" and x[0][0] == 0.25
    r["sink"](str(p))                                      # a second sink() without closing the first, then removal
    os.remove(p)
    r["summary"](Obj())
    r["sink"]()
    assert r["pdf"]("x.pdf") is None and r["dev.off"]() is None and r["plot"](Obj(), ask=False) is None
    np.testing.assert_allclose(r["sapply"](np.array([[1.0, 2.0], [3.0, 2.0]]), "sd"), [np.sqrt(2.0), 0.0])
    import pandas as pd
    const = pd.DataFrame({"a": np.full(1000, 0.1), "b": np.full(1000, 1.0 / 3.0), "c": np.arange(1000.0)})
    sd = pd.Series(r["sapply"](const, "sd"), index=const.columns)        # ABC.py:1778-1779
    assert list(sd[sd == 0].index) == ["a", "b"] and abs(sd["c"] - np.arange(1000.0).std(ddof=1)) < 1e-10


# ---- the same through the CUDA engine ---------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("method,statistic", [("rejection", "median"), ("loclinear", "median"), ("loclinear", "mode")])
def test_cv4abc_cuda_matches_oracle(engine, method, statistic):
    l0 = engine.k.launches
    _check_cv4abc(engine, method, statistic, N=3000, P=3, nval=5, tols=(0.05,))
    assert engine.k.launches > l0


@pytest.mark.gpu
def test_cv_facade_like_abcdls_calls_it(engine):
    """ABC.py:1900-1904 (per-column cv4abc + summary[0][0]), 886-892 (cv4postpr), 961-963 (gfit)."""
    import pandas as pd
    from abc_dls_b200 import rabc
    rabc.set_engine(engine)
    par, ss, tgt = synth.abc_tables(2000, 2, 2, "cpu")
    names = ["N_A", "T_B"]
    dfp, dfs = pd.DataFrame(par.T.numpy(), columns=names), pd.DataFrame(ss.T.numpy(), columns=names)
    samp = np.random.default_rng(2).choice(2000, 4, replace=False)
    for c in range(2):
        cvres = rabc.cv4abc(param=pd.DataFrame(dfp.iloc[:, c]), sumstat=pd.DataFrame(dfs.iloc[:, c]), nval=4, tols=0.05,
                            method="loclinear", trace=False, cvsamples=samp)
        want = O.cv4abc(par.numpy()[c], ss.numpy()[c], 4, 0.05, "loclinear", cvsamples=samp)
        s = cvres.summary(print_=False)
        assert abs(s[0][0] - O.summary_cv4abc(want)[0][0]) <= 1e-7 * O.summary_cv4abc(want)[0][0]
        assert cvres.names["parameter.names"] == [names[c]]
    joint = rabc.cv4abc(param=dfp, sumstat=dfs, nval=4, tols=0.05, method="rejection", cvsamples=samp)
    wj = O.cv4abc(par.numpy().T, ss.numpy().T, 4, 0.05, "rejection", cvsamples=samp)
    np.testing.assert_array_equal(joint.estim[0.05], wj.estim[0.05])       # type-7 medians of identical accepted sets
    assert pd.DataFrame(joint.summary(print_=False), columns=dfp.columns, index=[0.05]).shape == (1, 2)

    ids, ssp, _ = synth.postpr_tables(3000, "cpu")
    index = pd.Series(np.array(["BNDX", "MNDX", "SNDX"])[ids.numpy()])
    rng = np.random.default_rng(4)
    sp = np.concatenate([rng.choice(np.flatnonzero(ids.numpy() == k), 2, replace=False) for k in range(3)])
    cvm = rabc.cv4postpr(index=index, sumstat=pd.DataFrame(ssp.T.numpy()), nval=2, tol=0.05, method="rejection",
                         cvsamples=sp)
    wm = O.cv4postpr(index.to_numpy(), ssp.T.numpy(), 2, 0.05, cvsamples=sp)
    np.testing.assert_array_equal(cvm.estim[0.05], wm.estim[0.05])
    assert "Mean model posterior probabilities (rejection)" in cvm.summary_text()
    with pytest.raises(rabc.AbcError):
        rabc.cv4postpr(index=index, sumstat=pd.DataFrame(ssp.T.numpy()), nval=2, tol=0.05, method="mnlogistic")

    gs = np.random.default_rng(6).choice(2000, 5, replace=False)
    fit = rabc.gfit(pd.DataFrame(tgt.numpy()[None, :], columns=names), dfs, 5, tol=0.05, samples=gs)
    wg = O.gfit(tgt.numpy(), ss.numpy().T, 5, 0.05, samples=gs)
    np.testing.assert_allclose(fit.dist_sim, wg.dist_sim, rtol=1e-12)
    assert abs(fit.dist_obs - wg.dist_obs) <= 1e-12 * wg.dist_obs
    assert fit.summary(print_=False)["pvalue"] == O.summary_gfit(wg)["pvalue"]


@pytest.mark.gpu
def test_cv4abc_large_table_replays_the_graph_on_the_fast_path(engine):
    """N - 1 rows >= 2^20: the sampled-bracket front half runs on the hole buffer and the third validation on replays
    the CUDA graph of the second (same buffers, same shapes)."""
    N = (1 << 20) + 2
    par, ss, _ = synth.abc_tables(N, 2, 2, engine.device)
    samp = np.array([N - 1, 5, 700001, 0], dtype=np.int64)
    before = dict(engine.path_calls)
    got = CV.cv4abc(engine, par, ss, 4, 0.001, "rejection", cvsamples=samp)
    assert engine.path_calls["exact"] == before["exact"]                  # never the small-table path
    want = O.cv4abc(par.cpu().numpy().T, ss.cpu().numpy().T, 4, 0.001, "rejection", cvsamples=samp)
    np.testing.assert_array_equal(got.estim[0.001], want.estim[0.001])


# ---- one-CTA-per-fit path (csrc/small.cu): per-column mode on test-set sized tables ------------------------------------
def _columns(n, P, seed, decimals=None):
    par, ss, tgt = synth.abc_tables(n, P, P, "cpu")
    if decimals is not None:                       # exact ties in the statistic and in the distance
        scale = ss.abs().max(dim=1, keepdim=True).values
        ss = torch.round(ss / scale * 10 ** decimals) / 10 ** decimals * scale
    return par, ss, tgt


@pytest.mark.gpu
@pytest.mark.parametrize("n,decimals", [(3000, None), (10001, None), (4096, 3), (4096, 2)])
@pytest.mark.parametrize("method", ["rejection", "loclinear"])
def test_small_table_cv4abc_is_one_launch_and_matches_the_oracle(engine, n, decimals, method):
    if decimals == 2 and method == "loclinear":
        pytest.skip("two decimals leave one or two distinct statistics in the accepted region: lsfit is singular")
    P, nval, tol = 3, 7, 0.02
    par, ss, _ = _columns(n, P, 1, decimals)
    dpar, dss = par.to(engine.device), ss.to(engine.device)
    samp = np.random.default_rng(n).choice(n, nval, replace=False)
    samp[0], samp[1] = 0, n - 1
    for statistic in ("median", "mean"):
        l0 = engine.k.launches
        got = CV.cv4abc_columns(engine, dpar, dss, nval, tol, method, statistic=statistic, cvsamples=samp)
        assert engine.k.launches - l0 == 1                                  # P x nval fits, one launch
        for c in range(P):
            want = O.cv4abc(par.numpy()[c], ss.numpy()[c], nval, tol, method, statistic=statistic, cvsamples=samp)
            scale = np.abs(want.true).max()
            assert np.max(np.abs(got[c].estim[tol] - want.estim[tol])) <= 1e-9 * scale, (c, statistic)
            np.testing.assert_array_equal(got[c].true, want.true)
            one = CV.cv4abc(engine, dpar[c:c + 1], dss[c:c + 1], nval, [tol], method, statistic=statistic, cvsamples=samp)
            np.testing.assert_array_equal(one.estim[tol], got[c].estim[tol])   # per-column entry point: same kernel
    # row-major parameter table: same fits
    rm = dpar.t().contiguous().t()
    again = CV.cv4abc_columns(engine, rm, dss, nval, tol, method, statistic="mean", cvsamples=samp)
    for a, b in zip(again, got):
        np.testing.assert_array_equal(a.estim[tol], b.estim[tol])


@pytest.mark.gpu
def test_small_table_fit_equals_the_streaming_engine_and_reports_r_stops(engine):
    n, P, tol = 5000, 4, 0.01
    par, ss, tgt = _columns(n, P, 2)
    dpar, dss, dtgt = par.to(engine.device), ss.to(engine.device), tgt.to(engine.device)
    for method in ("rejection", "loclinear"):
        summ = engine.abc_columns_summary(dtgt, dpar, dss, tol, method)
        for c, r in enumerate(engine.abc_columns(dtgt, dpar, dss, tol, method)):
            o = O.abc(tgt.numpy()[c:c + 1], par.numpy()[c], ss.numpy()[c], tol, method)
            tab = O.summary_abc(o)
            scale = np.abs(par.numpy()[c]).max()
            assert abs(summ["min"][c] - tab[0, 0]) <= 1e-9 * scale and abs(summ["max"][c] - tab[6, 0]) <= 1e-9 * scale
            assert abs(summ["median"][c] - tab[2, 0]) <= 1e-9 * scale and abs(summ["mean"][c] - tab[3, 0]) <= 1e-9 * scale
            assert abs(summ["min"][c] - float(r.stats[0, 0])) <= 1e-9 * scale
    flat = dss.clone()
    flat[1] = 3.0                                                            # a constant statistic: R stops
    with pytest.raises(CV.AbcError, match="Zero variance in the summary statistics\\."):
        engine.abc_columns_summary(dtgt, dpar, flat, tol, "rejection")
    bad = dss.clone()
    bad[0, 17] = float("nan")
    with pytest.raises(CV.AbcError, match="non-finite"):
        engine.abc_columns_summary(dtgt, dpar, bad, tol, "loclinear")
    region = dss.clone()                                                     # accepted region of column 2 is constant
    region[2] = torch.where((region[2] - dtgt[2]).abs() < 0.2 * region[2].std(), dtgt[2], region[2])
    with pytest.raises(CV.AbcError, match="selected region"):
        engine.abc_columns_summary(dtgt, dpar, region, tol, "loclinear")
    big = torch.zeros((1, 200000), dtype=torch.float64, device=engine.device).normal_()
    l0 = engine.k.launches                                                   # too large for one CTA: streaming path
    engine.abc_columns_summary(big[:, 0], big, big, 0.001, "rejection")
    assert engine.k.launches - l0 > 1


@pytest.mark.gpu
def test_small_table_edge_cases_match_the_oracle(engine):
    """MAD == 0 (more than half of a column identical: R leaves it unscaled), nacc = 1, half the table accepted, even
    and odd row counts, target outside the simulated range."""
    rng = np.random.default_rng(12)
    for n in (1001, 1002):
        x = rng.normal(size=(3, n))
        x[1, : (6 * n) // 10] = 1.5                                     # column 1: MAD = 0
        x[2] = np.round(x[2], 1)                                        # column 2: heavy ties
        y = x + 0.1 * rng.normal(size=(3, n))
        tgt = np.array([0.3, 1.5001, 9.0])                             # column 2: target far outside
        dx, dy, dt = (torch.from_numpy(a).to(engine.device) for a in (x, y, tgt))
        for method in ("rejection", "loclinear"):
            for tol in (1.0 / n, 0.01, 0.5):
                if method == "loclinear" and tol * n < 8:
                    continue                                            # a fit needs a few points
                cols = [0, 2] if method == "loclinear" else [0, 1, 2]  # loclinear on the 60 % constant column: R stops
                got = engine.abc_columns_summary(dt[cols], dy[cols], dx[cols], tol, method)
                for i, c in enumerate(cols):
                    o = O.abc(tgt[c:c + 1], y[c], x[c], tol, method)
                    tab = O.summary_abc(o) if o.idx.size > 1 else None
                    v = o.adj_values if method == "loclinear" else o.unadj_values
                    scale = max(np.abs(y[c]).max(), np.abs(v).max())    # the far target extrapolates the fit by 1e7
                    assert abs(got["min"][i] - v.min()) <= 1e-9 * scale and abs(got["max"][i] - v.max()) <= 1e-9 * scale, \
                        (n, method, tol, c)
                    if tab is not None:
                        assert abs(got["median"][i] - tab[2, 0]) <= 1e-9 * scale, (n, method, tol, c)
                        assert abs(got["mean"][i] - tab[3, 0]) <= 1e-9 * scale, (n, method, tol, c)
