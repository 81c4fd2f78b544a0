"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (north_star): accepted index set bit-exact; dist / ds / weights / MAD bit-exact (fp64, no FMA
contraction, R's operation order); adjusted posterior within 1e-9 relative, where "relative" is against
max(|value|, scale of the parameter) because adjusted values may cross zero.
"""
import numpy as np
import pytest
import torch

from oracle import abc_oracle as O
from abc_dls_b200 import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-9


def _np(t):
    return t.detach().cpu().numpy()


def _rel(a, b):
    scale = np.maximum(np.abs(b).max(axis=0, keepdims=True), 1e-300)
    return float(np.max(np.abs(a - b) / scale))


def _check_abc(engine, par, ss, tgt, tol, method, hcorr=True):
    r = engine.abc(tgt, par, ss, tol, method, hcorr=hcorr, keep_dist=True)
    o = O.abc(_np(tgt), _np(par).T, _np(ss).T, tol, method, hcorr=hcorr)
    np.testing.assert_array_equal(_np(r.mad), o.mad)
    np.testing.assert_array_equal(_np(r.dist), o.dist)
    assert r.ds == o.ds and r.nacc == o.nacc
    np.testing.assert_array_equal(_np(r.idx), o.idx)            # bit-exact accepted set, row order
    np.testing.assert_array_equal(_np(r.unadj_values), o.unadj_values)
    np.testing.assert_array_equal(_np(r.ss), o.ss)
    if method == "loclinear":
        np.testing.assert_array_equal(_np(r.weights), o.weights)
        assert _rel(_np(r.adj_values), o.adj_values) < REL_TOL
        assert _rel(_np(r.pred)[None, :], o.pred[None, :]) < REL_TOL
        assert _rel(_np(r.coef), o.coef) < 1e-7
        assert _rel(_np(r.sigma2)[None, :], o.sigma2[None, :]) < 1e-9
        mn, mx = r.minmax()
        assert _rel(_np(mn)[None, :], o.adj_values.min(axis=0)[None, :]) < REL_TOL
        assert _rel(_np(mx)[None, :], o.adj_values.max(axis=0)[None, :]) < REL_TOL
    else:
        mn, mx = r.minmax()
        np.testing.assert_array_equal(_np(mn), o.unadj_values.min(axis=0))
        np.testing.assert_array_equal(_np(mx), o.unadj_values.max(axis=0))
    return r, o


@pytest.mark.parametrize("n", [2, 7, 1000, 100003, 1 << 20])
@pytest.mark.parametrize("d,P", [(1, 1), (3, 2), (16, 16)])
def test_rejection_parity(engine, n, d, P):
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    _check_abc(engine, par, ss, tgt, 0.3 if n < 100 else 0.01, "rejection")


@pytest.mark.parametrize("n,tol", [(1000, 0.1), (100003, 0.01), (1 << 20, 0.01)])
@pytest.mark.parametrize("d,P", [(1, 1), (3, 2), (10, 10), (16, 16)])
@pytest.mark.parametrize("hcorr", [True, False])
def test_loclinear_parity(engine, n, tol, d, P, hcorr):
    if n * tol < 4 * (d + 1):
        pytest.skip("fewer accepted rows than a well-posed regression needs")
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    _check_abc(engine, par, ss, tgt, tol, "loclinear", hcorr)


def test_odd_leading_dimension_and_unaligned_columns(engine):
    par, ss, tgt = synth.abc_tables(4097, 3, 3, engine.device)
    big = torch.zeros((3, 4099), dtype=torch.float64, device=engine.device)
    big[:, 1:4098] = ss
    _check_abc(engine, par, big[:, 1:4098], tgt, 0.05, "loclinear")


def test_ties_and_cap_rule(engine):
    """Heavily tied distances: integer-valued statistics -> many rows exactly at ds; R keeps the first
    nacc rows in row order among dist <= ds."""
    g = torch.Generator(device="cpu").manual_seed(3)
    n = 50000
    ss = torch.randint(0, 7, (2, n), generator=g).to(torch.float64).to(engine.device)
    par = torch.rand((2, n), generator=g, dtype=torch.float64).to(engine.device)
    tgt = torch.tensor([3.0, 3.0], dtype=torch.float64, device=engine.device)
    for tol in (0.001, 0.03, 0.5):
        r, o = _check_abc(engine, par, ss, tgt, tol, "rejection")
        assert (o.dist <= o.ds).sum() > o.nacc          # the cap rule actually bites
        assert r.idx.numel() == o.nacc


def test_zero_mad_column_is_left_unscaled(engine):
    g = torch.Generator(device="cpu").manual_seed(4)
    n = 20000
    ss = torch.rand((2, n), generator=g, dtype=torch.float64)
    ss[1] = torch.where(torch.rand(n, generator=g) < 0.7, torch.tensor(1.0, dtype=torch.float64), ss[1])
    par = torch.rand((1, n), generator=g, dtype=torch.float64)
    r, o = _check_abc(engine, par.to(engine.device), ss.to(engine.device),
                      torch.tensor([0.5, 0.9], dtype=torch.float64, device=engine.device), 0.02, "rejection")
    assert o.mad[1] == 0.0


def test_nacc_edges(engine):
    par, ss, tgt = synth.abc_tables(999, 2, 2, engine.device)
    r, o = _check_abc(engine, par, ss, tgt, 1.0, "rejection")       # everything accepted
    assert r.idx.numel() == 999
    r, o = _check_abc(engine, par, ss, tgt, 0.0005, "rejection")    # nacc == 1
    assert r.idx.numel() == 1


def test_per_column_mode_matches_p_separate_calls(engine):
    """ABC-DLS's rejection/loclinear path: P separate 1-D problems (ABC.py:1949-1953)."""
    par, ss, tgt = synth.abc_tables(10000, 19, 19, engine.device)      # cfg 1: SNDX schema, test_size rows
    out = engine.abc_columns(tgt, par, ss, 0.01, "loclinear")
    for c, r in enumerate(out):
        o = O.abc(_np(tgt)[c:c + 1], _np(par)[c], _np(ss)[c], 0.01, "loclinear")
        np.testing.assert_array_equal(_np(r.idx), o.idx)
        assert _rel(_np(r.adj_values), o.adj_values) < REL_TOL


def test_postpr_parity(engine):
    ids, ss, tgt = synth.postpr_tables(300000, engine.device)
    models = ["BNDX", "MNDX", "SNDX"]
    r = engine.postpr(tgt, ids, models, ss, 0.05, keep_dist=True)
    o = O.postpr(_np(tgt), np.array(models)[_np(ids)], _np(ss).T, 0.05)
    np.testing.assert_array_equal(_np(r.dist), o.dist)
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    np.testing.assert_array_equal(_np(r.counts), o.counts)
    assert r.ds == o.ds and (o.dist <= o.ds).sum() >= o.nacc
    np.testing.assert_allclose(_np(r.pred), o.pred, rtol=0, atol=0)


def test_errors_follow_r(engine):
    from abc_dls_b200.engine import AbcError
    par, ss, tgt = synth.abc_tables(500, 2, 2, engine.device)
    with pytest.raises(AbcError, match="Zero variance in the summary statistics\\."):
        engine.abc(tgt, par, torch.ones_like(ss), 0.1, "rejection")
    with pytest.raises(AbcError, match="Number of summary statistics"):
        engine.abc(tgt[:1], par, ss, 0.1, "rejection")
    bad = ss.clone()
    bad[0, 17] = float("nan")
    with pytest.raises(AbcError, match="non-finite"):
        engine.abc(tgt, par, bad, 0.1, "rejection")
    with pytest.raises(AbcError, match="neuralnet"):
        engine.abc(tgt, par, ss, 0.1, "neuralnet")


def test_smc_table_passes(engine):
    par, _, _ = synth.abc_tables(200001, 1, 12, engine.device)
    mn, mx = engine.oldrange(par)
    np.testing.assert_array_equal(_np(mn), _np(par).min(axis=1))
    np.testing.assert_array_equal(_np(mx), _np(par).max(axis=1))
    lo = mn + 0.1 * (mx - mn)
    hi = mx - 0.05 * (mx - mn)
    lo[3], hi[3] = mn[3], mx[3]
    idx = engine.box_filter(par, lo, hi)
    np.testing.assert_array_equal(_np(idx) + 2, O.narrowing_params(_np(par).T, _np(lo), _np(hi)))


def test_facade_accepts_pandas_like_the_reference_call_sites(engine):
    import pandas as pd
    from abc_dls_b200 import rabc
    rabc.set_engine(engine)
    par, ss, tgt = synth.abc_tables(10000, 3, 3, "cpu")
    names = ["N_A", "N_AF", "T_B"]
    dfp, dfs = pd.DataFrame(par.T.numpy(), columns=names), pd.DataFrame(ss.T.numpy(), columns=names)
    dft = pd.DataFrame(tgt.numpy()[None, :], columns=names)
    c = 1
    res = rabc.abc(target=pd.DataFrame(dft.iloc[:, c]), param=pd.DataFrame(dfp.iloc[:, c]),
                   sumstat=pd.DataFrame(dfs.iloc[:, c]), method="loclinear", tol=0.01)
    o = O.abc(tgt.numpy()[c:c + 1], par.numpy()[c], ss.numpy()[c], 0.01, "loclinear")
    tab = res.summary(print_=False)
    ot = O.summary_abc(o)
    assert abs(tab[0] - ot[0, 0]) <= REL_TOL * abs(ot[0, 0]) and abs(tab[6] - ot[6, 0]) <= REL_TOL * abs(ot[6, 0])
    np.testing.assert_allclose(np.array(tab)[[1, 2, 3, 5], 0], ot[[1, 2, 3, 5], 0], rtol=1e-9)
    assert res.summary_text().startswith("Data:")
    joint = rabc.abc(target=dft, param=dfp, sumstat=dfs, method="rejection", tol=0.01)
    oj = O.abc(tgt.numpy(), par.numpy().T, ss.numpy().T, 0.01, "rejection")
    np.testing.assert_array_equal(_np(joint.region_idx), oj.idx)
    np.testing.assert_allclose(np.array(joint.summary(print_=False)), O.summary_abc(oj), rtol=1e-9)
    assert not np.isnan(np.array(joint.summary(print_=False))).any()          # Mode row included
    # float32 frames (Keras predictions, ABC.py:1849) cross PCIe as float32 and are widened on the device: same result
    j32 = rabc.abc(target=dft, param=dfp.astype(np.float32), sumstat=dfs.astype(np.float32), method="rejection", tol=0.01)
    np.testing.assert_array_equal(_np(j32.region_idx), oj.idx)          # synth tables are float32-representable
    np.testing.assert_array_equal(_np(j32.unadj_values), oj.unadj_values)
    index = pd.Series(np.array(["BNDX", "MNDX", "SNDX"])[np.arange(30000) % 3])
    ids, ssp, tg = synth.postpr_tables(30000, "cpu")
    pp = rabc.postpr(target=pd.Series(tg.numpy()), index=index, sumstat=pd.DataFrame(ssp.T.numpy()), tol=0.05,
                     method="rejection")
    op = O.postpr(tg.numpy(), index.to_numpy(), ssp.T.numpy(), 0.05)
    np.testing.assert_array_equal(_np(pp.res.counts), op.counts)
    assert "Proportion of accepted simulations (rejection):" in pp.summary_text()


def test_param_table_may_stay_in_pinned_host_memory(engine):
    """Only the accepted rows of `param` are read: a pinned host table is gathered zero-copy."""
    par, ss, tgt = synth.abc_tables(1 << 20, 4, 3, engine.device)
    misses = engine.fast_misses
    ref = engine.abc(tgt, par, ss, 0.01, "loclinear")
    hpar = par.cpu().pin_memory()
    r = engine.abc(tgt, hpar, ss, 0.01, "loclinear")
    assert torch.equal(r.idx, ref.idx) and torch.equal(r.unadj_values, ref.unadj_values)
    assert torch.equal(r.adj_values, ref.adj_values) and engine.fast_misses == misses


def test_fast_path_is_taken_without_misses(engine):
    before = engine.fast_misses
    par, ss, tgt = synth.abc_tables((1 << 21) + 13, 16, 16, engine.device)
    r, o = _check_abc(engine, par, ss, tgt, 0.005, "loclinear")
    r2 = engine.abc(tgt, par, ss, 0.005, "loclinear")               # dist not kept: TMA kernel without the N-vector
    assert torch.equal(r2.idx, r.idx) and r2.dist is None
    assert engine.fast_misses == before


# ---- one-pass front half (csrc/fused.cuh): medians + MADs and the tolerance cut from ONE table read ----------------
def _check_fused(engine, par, ss, tgt, tol, method, expect="fused"):
    before = dict(engine.path_calls)
    misses = engine.fast_misses
    r = engine.abc(tgt, par, ss, tol, method)                       # dist not kept -> the one-pass path qualifies
    assert engine.path_calls[expect] == before[expect] + 1
    assert engine.fast_misses == misses
    o = O.abc(_np(tgt), _np(par).T, _np(ss).T, tol, method)
    np.testing.assert_array_equal(_np(r.mad), o.mad)
    assert r.ds == o.ds and r.nacc == o.nacc
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    np.testing.assert_array_equal(_np(r.unadj_values), o.unadj_values)
    np.testing.assert_array_equal(_np(r.ss), o.ss)
    np.testing.assert_array_equal(_np(r.dist_accepted), o.dist[o.idx])
    if method == "loclinear":
        np.testing.assert_array_equal(_np(r.weights), o.weights)
        assert _rel(_np(r.adj_values), o.adj_values) < REL_TOL
    return r, o


@pytest.mark.parametrize("n", [1 << 20, (1 << 21) + 14, (1 << 20) + 254])
@pytest.mark.parametrize("d,P", [(1, 1), (3, 2), (10, 10), (16, 16)])
def test_fused_one_pass_parity(engine, n, d, P):
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    _check_fused(engine, par, ss, tgt, 0.01, "loclinear" if d > 1 else "rejection")


def test_fused_odd_rows_in_an_even_pitch(engine):
    n = (1 << 20) + 77                                              # odd row count: ragged tile finished with plain stores
    par, ss, tgt = synth.abc_tables(n, 5, 4, engine.device)
    big = torch.zeros((5, n + 1), dtype=torch.float64, device=engine.device)
    big[:, :n] = ss
    _check_fused(engine, par, big[:, :n], tgt, 0.003, "loclinear")
    _check_fused(engine, par, ss, tgt, 0.003, "loclinear", expect="two_pass")   # odd pitch: TMA cannot stream it


def test_fused_small_and_large_tolerance(engine):
    par, ss, tgt = synth.abc_tables(1 << 21, 8, 8, engine.device)
    _check_fused(engine, par, ss, tgt, 0.0002, "loclinear")
    _check_fused(engine, par, ss, tgt, 0.2, "rejection")


def test_fused_miss_falls_back_and_stays_exact(engine):
    """Ties (integer statistics): the sampled MAD band collapses / MAD may be 0, the proof cannot hold -> the ladder
    falls back, the result is still R's."""
    g = torch.Generator(device="cpu").manual_seed(5)
    n = 1 << 20
    ss = torch.randint(0, 9, (3, n), generator=g).to(torch.float64).to(engine.device)
    par = torch.rand((2, n), generator=g, dtype=torch.float64).to(engine.device)
    tgt = torch.tensor([4.0, 3.0, 5.0], dtype=torch.float64, device=engine.device)
    r = engine.abc(tgt, par, ss, 0.01, "rejection")
    o = O.abc(_np(tgt), _np(par).T, _np(ss).T, 0.01, "rejection")
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    assert r.ds == o.ds


def test_fused_postpr(engine):
    ids, ss, tgt = synth.postpr_tables(3000000, engine.device)
    models = ["BNDX", "MNDX", "SNDX"]
    r = engine.postpr(tgt, ids, models, ss, 0.05)
    o = O.postpr(_np(tgt), np.array(models)[_np(ids)], _np(ss).T, 0.05)
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    np.testing.assert_array_equal(_np(r.counts), o.counts)
    assert r.ds == o.ds


def test_cuda_graph_replay_matches_eager_and_tracks_the_target(engine):
    """Call 1 runs eagerly, call 2 captures the device work into a CUDA graph, later calls replay it.  The replayed
    results must equal the eager ones bit for bit, follow a changed target (copied into the graph's buffer), and stay
    valid after the next call (results are copied out of the graph's buffers)."""
    par, ss, tgt = synth.abc_tables((1 << 20) + 64, 6, 5, engine.device)
    engine._graphs.clear(); engine._seen.clear()
    r1 = engine.abc(tgt, par, ss, 0.01, "loclinear")
    r2 = engine.abc(tgt, par, ss, 0.01, "loclinear")
    assert len(engine._graphs) == 1
    r3 = engine.abc(tgt, par, ss, 0.01, "loclinear")
    for r in (r2, r3):
        assert torch.equal(r.idx, r1.idx) and torch.equal(r.adj_values, r1.adj_values) and r.ds == r1.ds
        assert torch.equal(r.stats, r1.stats) and r.aic == r1.aic
    keep = r3.adj_values.clone()
    tgt2 = tgt * 1.01
    r4 = engine.abc(tgt2, par, ss, 0.01, "loclinear")                 # same key, new target values: replay
    assert torch.equal(r3.adj_values, keep)                           # earlier results were not overwritten
    o = O.abc(_np(tgt2), _np(par).T, _np(ss).T, 0.01, "loclinear")
    np.testing.assert_array_equal(_np(r4.idx), o.idx)
    assert r4.ds == o.ds and _rel(_np(r4.adj_values), o.adj_values) < REL_TOL
    assert abs(r4.aic - o.aic) <= 1e-9 * abs(o.aic)
    # rarely read arrays are copied out of the graph's buffers on first access: fine for the latest call, an error
    # (never the newer call's data) once the buffers have been reused
    np.testing.assert_array_equal(_np(r4.ss), o.ss)
    np.testing.assert_array_equal(_np(r4.unadj_values), o.unadj_values)
    with pytest.raises(RuntimeError, match="reused the graph's buffers"):
        r3.ss


# ---- BASELINE.json full sizes ------------------------------------------------------------------------------------------
def test_cfg3_full_size_against_the_oracle(engine):
    """BASELINE config 3: 1e7 simulations x 10 statistics x 10 parameters, loclinear, tol 0.001, one GPU."""
    par, ss, tgt = synth.abc_tables(10_000_000, 10, 10, engine.device)
    r, o = _check_fused(engine, par, ss, tgt, 0.001, "loclinear")
    assert r.nacc == 10_000 and r.idx.numel() == 10_000
    mn, mx = r.minmax()
    assert _rel(_np(mn)[None, :], o.adj_values.min(axis=0)[None, :]) < REL_TOL
    assert _rel(_np(mx)[None, :], o.adj_values.max(axis=0)[None, :]) < REL_TOL


def test_cfg5_full_size_three_independent_paths_agree(engine):
    """BASELINE config 5 at its full size (1e8 x 16 x 16, tol 0.01): too big for the CPU oracle in a test, so the three
    front halves - one table read with interval reasoning, two reads with sampled brackets, thirteen reads with the
    distribution-independent radix select - must give bit-identical MADs, ds and accepted rows, and the accepted set
    must satisfy R's definition against the full distance vector (size-independent properties)."""
    N = 100_000_000
    par, ss, tgt = synth.abc_tables(N, 16, 16, engine.device)
    engine._graphs.clear()
    before, misses = dict(engine.path_calls), engine.fast_misses
    a = engine.abc(tgt, par, ss, 0.01, "loclinear")                              # one-pass
    b = engine.abc(tgt, par, ss, 0.01, "loclinear", keep_dist=True)              # two-pass (returns dist)
    c = engine.abc(tgt, par, ss, 0.01, "loclinear", keep_dist=True, _level=2)    # radix select
    assert engine.path_calls["fused"] == before["fused"] + 1
    assert engine.path_calls["two_pass"] == before["two_pass"] + 1
    assert engine.path_calls["exact"] == before["exact"] + 1 and engine.fast_misses == misses
    for r in (b, c):
        assert torch.equal(r.mad, a.mad) and r.ds == a.ds and torch.equal(r.idx, a.idx)
        assert torch.equal(r.unadj_values, a.unadj_values) and torch.equal(r.weights, a.weights)
        assert _rel(_np(r.adj_values), _np(a.adj_values)) < 1e-12
    assert torch.equal(b.dist, c.dist)
    # R's definition on the full vector: nacc = ceil(N tol); ds = nacc-th smallest; first nacc rows with dist <= ds
    nacc = 1_000_000
    assert a.nacc == nacc and a.idx.numel() == nacc
    assert int((b.dist < a.ds).sum()) < nacc <= int((b.dist <= a.ds).sum())
    assert bool((a.idx[1:] > a.idx[:-1]).all())
    assert torch.equal(b.dist[a.idx], a.dist_accepted) and bool((a.dist_accepted <= a.ds).all())
    first = torch.nonzero(b.dist <= a.ds).view(-1)[:nacc]
    assert torch.equal(first, a.idx)
    del b, c


@pytest.mark.parametrize("d,P", [(2, 2), (7, 3), (15, 4)])
def test_fused_odd_column_counts(engine, d, P):
    """Column counts that do not fill the 16-warp CTA / the paired 16-byte emission stores."""
    par, ss, tgt = synth.abc_tables((1 << 20) + 512, d, P, engine.device)
    _check_fused(engine, par, ss, tgt, 0.02, "loclinear")


def test_fused_tiny_tolerance_and_half_the_table(engine):
    """nacc = 21 (the sampled threshold is far above ds: many more candidates than needed, still exact) and
    tol = 0.5 (every second row is accepted: the candidate list is the whole table)."""
    par, ss, tgt = synth.abc_tables(1 << 21, 4, 2, engine.device)
    misses = engine.fast_misses
    r = engine.abc(tgt, par, ss, 1e-5, "rejection")
    o = O.abc(_np(tgt), _np(par).T, _np(ss).T, 1e-5, "rejection")
    assert o.nacc == 21 and r.ds == o.ds
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    r = engine.abc(tgt, par, ss, 0.5, "rejection")
    o = O.abc(_np(tgt), _np(par).T, _np(ss).T, 0.5, "rejection")
    assert r.ds == o.ds
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    assert engine.fast_misses - misses <= 2           # a miss is allowed here (ladder), a wrong answer is not


def test_fused_target_far_from_the_data(engine):
    """Large offset between target and data relative to the spread: the rounding slack eta of the proof must cover
    R's cancellation in (x/m - t/m)."""
    g = torch.Generator(device="cpu").manual_seed(11)
    n = 1 << 20
    ss = (1.0e6 + torch.randn((3, n), generator=g, dtype=torch.float64)).to(engine.device)
    par = torch.rand((2, n), generator=g, dtype=torch.float64).to(engine.device)
    tgt = torch.tensor([1.0e6 + 0.1, 1.0e6 - 0.2, 1.0e6], dtype=torch.float64, device=engine.device)
    r = engine.abc(tgt, par, ss, 0.01, "loclinear")
    o = O.abc(_np(tgt), _np(par).T, _np(ss).T, 0.01, "loclinear")
    np.testing.assert_array_equal(_np(r.idx), o.idx)
    assert r.ds == o.ds and _rel(_np(r.adj_values), o.adj_values) < REL_TOL


@pytest.mark.parametrize("n,d,P,tol,method", [(50000, 3, 5, 0.02, "loclinear"), ((1 << 20) + 77, 4, 16, 0.01, "loclinear"),
                                              ((1 << 20) + 77, 4, 3, 0.003, "rejection")])
def test_row_major_param_table_gives_the_same_result(engine, n, d, P, tol, method):
    """`param` as numpy / pandas hand it over (n x P C order, passed as its (P, n) transposed view), on the device and
    in pinned host memory: same accepted rows, same values as the column-major table (exact and one-pass paths)."""
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    ref = engine.abc(tgt, par, ss, tol, method)
    rm = par.t().contiguous()                              # (n, P) row-major
    assert rm.t().stride() == (1, P)
    host = torch.empty((n, P), dtype=torch.float64, pin_memory=True)
    host.copy_(rm)
    for table in (rm.t(), host.t()):
        for _ in range(3):                                 # eager, capture, replay
            r = engine.abc(tgt, table, ss, tol, method)
            assert torch.equal(r.idx, ref.idx) and torch.equal(r.unadj_values, ref.unadj_values)
            assert torch.equal(r.values, ref.values) and torch.equal(r.stats, ref.stats)
    cols = engine.abc_columns(tgt[:min(d, P)], rm.t()[:min(d, P)], ss[:min(d, P)], tol, "rejection")
    cref = engine.abc_columns(tgt[:min(d, P)], par[:min(d, P)], ss[:min(d, P)], tol, "rejection")
    for a, b in zip(cols, cref):
        assert torch.equal(a.idx, b.idx) and torch.equal(a.unadj_values, b.unadj_values)


def test_cfg4_smc_round_per_column_on_a_large_table(engine):
    """BASELINE cfg4 shape at 2^21 rows per round: 12 parameters, per-column rejection (ABC.py:2805-2840) -> posterior
    min / max -> updating_newrange + noise_injection_update (decrease .95, increase .01, hard range) -> box filter of
    the simulated parameters -> lmrd; two rounds, the second on the rows the first one kept."""
    from abc_dls_b200 import smc
    n, P, tol = 1 << 21, 12, 0.01
    par, ss, tgt = synth.abc_tables(n, P, P, engine.device)
    hard_min, hard_max = _np(par).min(axis=1) - 1.0, _np(par).max(axis=1) + 1.0
    for rnd in range(2):
        res = engine.abc_columns(tgt, par, ss, tol, "rejection")
        pmin = np.array([float(r.minmax()[0][0]) for r in res])
        pmax = np.array([float(r.minmax()[1][0]) for r in res])
        omin, omax = engine.oldrange(par)
        rng = smc.updating_newrange(pmin, pmax, _np(omin), _np(omax), 0.95)
        rng = smc.noise_injection_update(rng, _np(omin), _np(omax), 0.01, hard_min, hard_max, 0.95)
        keep = engine.box_filter(par, torch.from_numpy(rng.min), torch.from_numpy(rng.max))
        want = O.smc_round(_np(tgt), _np(par).T, _np(ss).T, tol, "rejection", _np(par).T, 0.95, 0.01, hard_min, hard_max)
        np.testing.assert_array_equal(pmin, want["post_min"])
        np.testing.assert_array_equal(pmax, want["post_max"])
        np.testing.assert_array_equal(rng.min, want["min"])
        np.testing.assert_array_equal(rng.max, want["max"])
        np.testing.assert_array_equal(smc.csv_linenumbers(keep), want["lines"])
        assert smc.lmrd_calculation(rng, hard_min, hard_max) == want["lmrd"]
        assert 0 < keep.numel() < par.shape[1]
        par, ss = par[:, keep].contiguous(), ss[:, keep].contiguous()      # next round: the narrowed table


@pytest.mark.parametrize("n", [(1 << 18) + 6, 300_002, 1_000_000])
@pytest.mark.parametrize("d,P,tol", [(16, 16, 0.01), (3, 2, 0.001), (1, 1, 0.05)])
def test_fused_one_pass_on_tables_just_above_the_threshold(engine, n, d, P, tol):
    """FAST_MIN_ROWS = 2^18 rows per rank: the sampled brackets (8 sample CTAs per column at this size) must still
    prove themselves - exact results, no misses - on the smallest tables that take the one-pass path (the 1e6-row point
    of the BASELINE cfg5 sweep among them)."""
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    _check_fused(engine, par, ss, tgt, tol, "loclinear" if d > 1 else "rejection")


# ---- 16 < d <= 32: 128-row tiles, two columns per consumer warp (csrc/fused_wide.cuh) -----------------------------------
@pytest.mark.parametrize("n,d,P", [((1 << 20) + 130, 17, 5), ((1 << 20) + 64, 19, 19), ((1 << 21) + 8, 24, 24),
                                   (1 << 20, 32, 32), ((1 << 20) + 254, 31, 3)])
def test_fused_one_pass_wide_tables(engine, n, d, P):
    """The reference's models have 19 and 24 parameters (tests/Model.info): a joint call then has 19 / 24 statistics
    and must still take ONE table read."""
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    _check_fused(engine, par, ss, tgt, 0.01, "loclinear")


def test_fused_wide_with_the_second_stage_sample_and_ties(engine):
    n = (1 << 23) + 512                                             # fine sample on (engine.FINE_MIN_ROWS)
    par, ss, tgt = synth.abc_tables(n, 24, 4, engine.device)
    ss[3] = torch.round(ss[3], decimals=2)                          # one coarse column: ties inside the bands
    _check_fused(engine, par, ss, tgt, 0.002, "rejection")
