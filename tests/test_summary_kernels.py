"""summary.abc from this repo's kernels only (csrc/summary.cu + csrc/select.cu: weighted radix select, type-7 ranks, density
binning + direct gaussian convolution) against the oracle's restatement of R (quantile type 7, quantreg::rq weighted
order statistics, stats::density / bw.nrd0 mode) and against the framework-op formulation it replaces."""
import numpy as np
import pytest
import torch

from oracle import abc_oracle as O
from abc_dls_b200 import summary as S

pytestmark = pytest.mark.gpu


def _case(n, P, seed, ties=False):
    rng = np.random.default_rng(seed)
    vals = rng.normal(size=(n, P)) * (10.0 ** rng.integers(-3, 4, P)) + rng.normal(size=P) * 50.0
    if P > 1:
        vals[:, 1] = np.concatenate([rng.normal(0.0, 1.0, n - n // 4), rng.normal(6.0, 0.5, n // 4)])   # bimodal
    if ties:
        vals = np.round(vals, 1)
    w = rng.random(n)
    w[rng.integers(0, n, max(1, n // 50))] = 0.0           # the farthest accepted row has weight 0 in R, too
    return vals, w


@pytest.mark.parametrize("n,P,ties", [(4000, 3, False), (200_000, 5, False), (50_000, 2, True), (9, 2, False), (2, 1, False)])
@pytest.mark.parametrize("method", ["rejection", "loclinear"])
def test_device_summary_matches_the_oracle(engine, n, P, ties, method):
    vals, w = _case(n, P, 11 + n, ties)
    res = O.AbcOracleResult.__new__(O.AbcOracleResult)
    res.method, res.unadj_values, res.adj_values, res.weights = method, vals, vals, w
    want = O.summary_abc(res)
    ww = w if method == "loclinear" else np.ones(n)
    stats = np.stack([vals.min(0), vals.max(0), vals.sum(0), (ww[:, None] * vals).sum(0),
                      (ww[:, None] * vals * vals).sum(0), np.full(P, ww.sum())], axis=1)
    dev = engine.device
    tv, tw, ts = torch.from_numpy(vals).to(dev), torch.from_numpy(w).to(dev), torch.from_numpy(stats).to(dev)
    got = np.asarray(S.summary_table(tv, tw, ts, weighted=(method == "loclinear"), eng=engine))
    ref = np.asarray(S.summary_table(tv, tw, ts, weighted=(method == "loclinear")))      # framework ops (sort / fft)
    scale = np.abs(want).max(axis=0)
    assert np.all(np.abs(got - want) <= 1e-9 * scale), (got - want) / scale
    assert np.all(np.abs(got - ref) <= 1e-9 * scale)
    assert not np.isnan(got).any()
    # the percentile rows are sample values (weighted) / two-point interpolations (type 7): identical, not just close
    rows = [1, 2, 5]
    np.testing.assert_array_equal(got[rows], want[rows])


def test_device_summary_is_what_the_facade_prints(engine):
    """rabc.AbcFit.summary() goes through the kernels on a CUDA engine; same text as the oracle's table."""
    from abc_dls_b200 import rabc, synth
    par, ss, tgt = synth.abc_tables(60_000, 3, 3, engine.device)
    r = engine.abc(tgt, par, ss, 0.05, "loclinear")
    fit = rabc.AbcFit(r, engine, ["a", "b", "c"], ["s1", "s2", "s3"])
    tab = np.asarray(fit.summary(print_=False))
    o = O.abc(tgt.cpu().numpy(), par.cpu().numpy().T, ss.cpu().numpy().T, 0.05, "loclinear")
    want = O.summary_abc(o)
    assert np.all(np.abs(tab - want) <= 1e-9 * np.abs(want).max(axis=0))
