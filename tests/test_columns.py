"""Per-column mode as ONE pass (csrc/columns.cuh) against the oracle's P separate 1-D abc() calls and against the
engine's own loop over abc(): min / max of the accepted parameters bit-exact, ds and MAD bit-exact, mean 1e-12."""
import numpy as np
import pytest
import torch

from oracle import abc_oracle as O
from abc_dls_b200 import synth

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def _oracle_columns(tgt, par, ss, tol):
    P = par.shape[1]
    out = {k: np.empty(P) for k in ("min", "max", "mean", "median", "ds", "mad")}
    for c in range(P):
        o = O.abc(tgt[c:c + 1], par[:, c], ss[:, c:c + 1], tol, "rejection")
        v = o.unadj_values[:, 0]
        out["min"][c], out["max"][c], out["mean"][c] = v.min(), v.max(), O.r_mean(v)
        out["median"][c], out["ds"][c], out["mad"][c] = O.r_median(v), o.ds, o.mad[0]
        assert v.size == o.nacc
    return out


def _check(engine, par, ss, tgt, tol, want=("min", "max", "mean", "median")):
    before = engine.path_calls.get("columns", 0)
    got = engine.abc_columns_summary(tgt, par, ss, tol, "rejection", want=want)
    assert engine.path_calls.get("columns", 0) == before + 1, "the one-pass per-column path was not taken"
    ref = _oracle_columns(_np(tgt), _np(par).T, _np(ss).T, tol)
    for k in ("min", "max", "ds", "mad"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    if "mean" in want:
        np.testing.assert_allclose(got["mean"], ref["mean"], rtol=1e-12)
    if "median" in want:
        np.testing.assert_array_equal(got["median"], ref["median"])
    return got


@pytest.mark.parametrize("n,P,tol,layout", [((1 << 20) + 4097, 5, 0.01, "col"), ((1 << 21) + 1, 3, 0.002, "row"),
                                            ((1 << 20) + 64, 1, 0.05, "col"), ((1 << 18) + 1, 4, 0.01, "row")])
def test_columns_one_pass_matches_the_oracle(engine, n, P, tol, layout):
    par, ss, tgt = synth.abc_tables(n, P, P, engine.device)
    if layout == "row":
        par = par.t().contiguous().t()
    _check(engine, par, ss, tgt, tol)


def test_columns_ties_at_the_threshold_follow_the_cap_rule(engine):
    """Statistics rounded to a coarse grid: thousands of rows share every distance, ds falls inside a block of ties and
    R's cumsum cap rule takes the first of them in row order - with DIFFERENT parameter values behind equal statistics."""
    n, P = (1 << 20) + 300, 3
    par, ss, tgt = synth.abc_tables(n, P, P, engine.device)
    span = ss.max(dim=1).values - ss.min(dim=1).values
    ss = (torch.round(ss / span[:, None] * 3000.0) * span[:, None] / 3000.0).contiguous()   # ~350 rows per grid value
    ss = ss.to(torch.float32).to(torch.float64)
    got = _check(engine, par, ss, tgt, 0.01)
    assert np.all(np.isfinite(got["min"]))


def test_columns_matches_the_loop_over_abc_and_repeated_calls_replay_the_graph(engine):
    n, P, tol = (1 << 22) + 7, 12, 0.01
    par, ss, tgt = synth.abc_tables(n, P, P, engine.device)
    a = [engine.abc_columns_summary(tgt, par, ss, tol, "rejection", want=("min", "max", "mean")) for _ in range(3)]
    engine.use_columns = False
    try:
        b = engine.abc_columns_summary(tgt, par, ss, tol, "rejection", want=("min", "max", "mean"))
    finally:
        engine.use_columns = True
    for k in ("min", "max"):
        np.testing.assert_array_equal(a[0][k], b[k])
        np.testing.assert_array_equal(a[2][k], b[k])
    np.testing.assert_allclose(a[2]["mean"], b["mean"], rtol=1e-12)
    # another target with the same tables: the replayed graph must follow it
    tgt2 = (tgt * 1.01).contiguous()
    c = engine.abc_columns_summary(tgt2, par, ss, tol, "rejection", want=("min", "max"))
    engine.use_columns = False
    try:
        d = engine.abc_columns_summary(tgt2, par, ss, tol, "rejection", want=("min", "max"))
    finally:
        engine.use_columns = True
    for k in ("min", "max"):
        np.testing.assert_array_equal(c[k], d[k])


def test_columns_target_outside_the_data_and_constant_column_fall_back(engine):
    n, P = (1 << 20) + 11, 2
    par, ss, tgt = synth.abc_tables(n, P, P, engine.device)
    far = tgt.clone()
    far[0] = ss[0].max() + 10.0 * (ss[0].max() - ss[0].min())          # every distance on one side of the target
    _check(engine, par, ss, far, 0.01, want=("min", "max"))
    ss2 = ss.clone()
    ss2[1] = 3.25                                                      # MAD 0 and zero variance: R stops -> exact path
    from abc_dls_b200.engine import AbcError
    with pytest.raises(AbcError, match="Zero variance"):
        engine.abc_columns_summary(tgt, par, ss2, 0.01, "rejection", want=("min", "max"))
