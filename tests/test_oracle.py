"""Known-answer and self-consistency tests of the CPU oracle (oracle/abc_oracle.py).

The reference pins nothing on this path (tests/test_ABCDLS.py only checks that files exist) and R is not
available, so the oracle is checked against hand-computed cases, against a second independent
formulation of each step, and against invariances of the algorithm.
"""
import math

import numpy as np
import pytest

from oracle import abc_oracle as O


def test_median_and_mad_known_answers():
    assert O.r_median(np.array([3.0, 1.0, 2.0])) == 2.0
    assert O.r_median(np.array([4.0, 1.0, 3.0, 2.0])) == 2.5
    assert O.r_median(np.array([5.0])) == 5.0
    # R: mad(c(1,2,3,4,100)) = 1.4826 * median(|x-3|) = 1.4826 * 1
    assert O.r_mad(np.array([1.0, 2.0, 3.0, 4.0, 100.0])) == 1.4826
    # R: mad(c(1,2,3,4)) -> median 2.5, |x-2.5| = 1.5,.5,.5,1.5 -> median 1.0
    assert O.r_mad(np.array([1.0, 2.0, 3.0, 4.0])) == 1.4826
    assert O.r_mad(np.array([7.0, 7.0, 7.0, 9.0])) == 0.0


def test_median_matches_numpy_on_random_columns():
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 11, 1000, 1001):
        x = rng.normal(size=n) * 10.0 ** rng.integers(-3, 4)
        assert O.r_median(x) == np.median(x)


def test_quantile_type7_known():
    x = np.arange(1, 11, dtype=float)
    np.testing.assert_allclose(O.r_quantile7(x, [0.025, 0.5, 0.975]), [1.225, 5.5, 9.775], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(O.r_quantile7(x, [0.0, 1.0]), [1.0, 10.0])


def test_nacc_is_ceiling_of_double_product():
    d = np.arange(100, dtype=float)
    assert O.tolerance_cut(d, 0.01)[0] == 1
    assert O.tolerance_cut(d, 0.011)[0] == 2
    assert O.tolerance_cut(np.arange(10_000_000, dtype=float), 0.001)[0] == 10000
    assert O.tolerance_cut(np.arange(3_000_000, dtype=float), 0.05)[0] == 150000


def test_cap_rule_first_nacc_rows_in_row_order():
    # 10 rows, tol .3 -> nacc 3; sorted dist = 0,1,1,1,... -> ds = 1; four rows have dist <= ds;
    # R keeps the first three IN ROW ORDER, dropping row 7 (dist 0 < ds!) because it comes last.
    dist = np.array([1.0, 5.0, 1.0, 9.0, 1.0, 8.0, 7.0, 0.0, 6.0, 4.0])
    nacc, ds, wt1 = O.tolerance_cut(dist, 0.3)
    assert (nacc, ds) == (3, 1.0)
    assert np.flatnonzero(wt1).tolist() == [0, 2, 4]
    _, _, wt_nocap = O.tolerance_cut(dist, 0.3, cap_rule="none")
    assert np.flatnonzero(wt_nocap).tolist() == [0, 2, 4, 7]


def test_distance_hand_case_and_zero_mad_column():
    ss = np.array([[1.0, 5.0], [2.0, 5.0], [3.0, 5.0], [4.0, 6.0], [100.0, 5.0]])
    tgt = np.array([2.0, 5.5])
    scaled, t, mad, dist = O.scale_and_distance(tgt, ss)
    assert mad[0] == 1.4826 and mad[1] == 0.0        # second column: mad 0 -> left unscaled
    np.testing.assert_array_equal(scaled[:, 1], ss[:, 1])
    assert t[1] == 5.5
    exp = np.sqrt((ss[:, 0] / 1.4826 - 2.0 / 1.4826) ** 2 + (ss[:, 1] - 5.5) ** 2)
    np.testing.assert_array_equal(dist, exp)


def _tables(n, d, P, seed):
    rng = np.random.default_rng(seed)
    lo, hi = rng.uniform(0, 10, P), rng.uniform(20, 1000, P)
    theta = (lo + (hi - lo) * rng.uniform(size=(n, P))).astype(np.float32).astype(np.float64)
    ss = (theta[:, np.arange(d) % P] + 0.1 * (hi - lo)[np.arange(d) % P] * rng.normal(size=(n, d))).astype(
        np.float32).astype(np.float64)
    target = (lo + 0.37 * (hi - lo))[np.arange(d) % P]
    return theta, ss, target


def test_rejection_is_sort_based_selection():
    theta, ss, tgt = _tables(5000, 4, 3, 1)
    r = O.abc(tgt, theta, ss, 0.02, "rejection")
    order = np.argsort(r.dist, kind="stable")
    assert r.nacc == 100 and r.idx.size == 100
    assert set(r.idx.tolist()) == set(order[:100].tolist())   # continuous data: no ties
    assert np.all(np.diff(r.idx) > 0)
    np.testing.assert_array_equal(r.unadj_values, theta[r.idx])
    assert r.ds == np.sort(r.dist)[99]


@pytest.mark.parametrize("d,P", [(1, 1), (3, 2), (10, 10)])
def test_loclinear_matches_centred_normal_equations(d, P):
    """Second formulation: weighted normal equations centred on the target, solved by Cholesky.  The
    intercept of the centred fit is the prediction at the target."""
    theta, ss, tgt = _tables(20000, d, P, 2 + d)
    r = O.abc(tgt, theta, ss, 0.05, "loclinear")
    X = (r.ss / r.mad) - (tgt / r.mad)
    A = np.column_stack([np.ones(len(X)), X])
    G = (A * r.weights[:, None]).T @ A
    beta = np.linalg.solve(G, (A * r.weights[:, None]).T @ r.unadj_values)
    res = r.unadj_values - A @ beta
    mu = res.mean(axis=0)
    res = res - mu
    gam = np.linalg.solve(G, (A * r.weights[:, None]).T @ np.log(res ** 2))
    adj = (beta[0] + mu) + np.sqrt(np.exp(gam[0])) * res / np.sqrt(np.exp(A @ gam))
    scale = np.abs(r.adj_values).max(axis=0)
    assert np.max(np.abs(adj - r.adj_values) / scale) < 1e-10
    # farthest accepted row has weight exactly 0 (bandwidth = ds)
    assert r.weights.min() == 0.0 and r.weights.max() <= 1.0


def test_invariances():
    theta, ss, tgt = _tables(4000, 3, 3, 7)
    base = O.abc(tgt, theta, ss, 0.05, "loclinear")
    # scaling a summary-statistic column (and its target) by a power of two changes nothing
    ss2, tgt2 = ss.copy(), tgt.copy()
    ss2[:, 1] *= 8.0
    tgt2[1] *= 8.0
    r2 = O.abc(tgt2, theta, ss2, 0.05, "loclinear")
    np.testing.assert_array_equal(base.idx, r2.idx)
    np.testing.assert_array_equal(base.dist, r2.dist)
    np.testing.assert_allclose(base.adj_values, r2.adj_values, rtol=1e-12)
    # tol = 1 accepts everything
    assert O.abc(tgt, theta, ss, 1.0, "rejection").idx.size == 4000
    # permuting rows (no ties) permutes the accepted set
    perm = np.random.default_rng(3).permutation(4000)
    rp = O.abc(tgt, theta[perm], ss[perm], 0.05, "rejection")
    assert set(perm[rp.idx].tolist()) == set(base.idx.tolist())


def test_duplicated_table_doubles_postpr_counts():
    rng = np.random.default_rng(5)
    n = 3000
    ids = np.array(["BNDX", "MNDX", "SNDX"])[np.arange(n) % 3]
    logits = rng.normal(size=(n, 3)) + 2.0 * (np.arange(3)[None, :] == (np.arange(n) % 3)[:, None])
    sm = np.round(np.exp(logits) / np.exp(logits).sum(1, keepdims=True), 5)
    a = O.postpr([0.2, 0.5, 0.3], ids, sm, 0.1)
    assert a.models == ["BNDX", "MNDX", "SNDX"] and a.counts.sum() == 300
    b = O.postpr([0.2, 0.5, 0.3], np.concatenate([ids, ids]), np.vstack([sm, sm]), 0.1)
    assert b.counts.sum() == 600


def test_errors_follow_r():
    theta, ss, tgt = _tables(100, 2, 2, 9)
    with pytest.raises(O.AbcError, match="Zero variance in the summary statistics\\."):
        O.abc(tgt, theta, np.ones_like(ss), 0.1, "rejection")
    with pytest.raises(O.AbcError, match="Number of summary statistics"):
        O.abc(tgt[:1], theta, ss, 0.1, "rejection")
    const_region = ss.copy()
    const_region[:, 0] = np.where(np.arange(100) < 50, 1.0, np.arange(100))
    const_region[:, 1] = np.where(np.arange(100) < 50, 2.0, 7.0 + np.arange(100))
    with pytest.raises(O.AbcError, match="selected region"):
        O.abc(np.array([1.0, 2.0]), theta, const_region, 0.1, "loclinear")
    with pytest.raises(O.AbcError, match="At least two"):
        O.postpr([1.0, 2.0], ["a"] * 100, ss, 0.1)


def test_summary_rows():
    theta, ss, tgt = _tables(2000, 1, 1, 11)
    rej = O.abc(tgt, theta, ss, 0.1, "rejection")
    tab = O.summary_abc(rej)
    x = rej.unadj_values[:, 0]
    assert tab[0, 0] == x.min() and tab[6, 0] == x.max()
    assert tab[2, 0] == np.quantile(x, 0.5) and math.isclose(tab[3, 0], x.mean(), rel_tol=1e-15)
    ll = O.abc(tgt, theta, ss, 0.1, "loclinear")
    tl = O.summary_abc(ll)
    assert tl[0, 0] == ll.adj_values.min() and tl[6, 0] == ll.adj_values.max()
    assert tl[1, 0] <= tl[2, 0] <= tl[5, 0]


# ---- SMC range rules: pinned on the reference's own functions (tests/test_reference_callsites.py part B: the oracle
# and abc_dls_b200.smc both reproduce tests/golden/ref_smc_rules.npz, generated by running ABC.py:2881-3028 here) -----
def test_narrowing_and_lmrd():
    rng = np.random.default_rng(4)
    par = rng.uniform(0, 1, size=(1000, 3))
    lines = O.narrowing_params(par, [0.1, 0.2, 0.0], [0.9, 0.8, 1.0])
    keep = (par[:, 0] >= .1) & (par[:, 0] <= .9) & (par[:, 1] >= .2) & (par[:, 1] <= .8)
    np.testing.assert_array_equal(lines, np.flatnonzero(keep) + 2)
    assert math.isclose(O.lmrd_calculation([0, 0], [1, 2], [0, 0], [2, 4]), math.log(0.5))


# ---- summary.abc Mode row: R density() restated (stats::density.default, bw.nrd0, C_BinDist, FFT convolution) --------
def test_bw_nrd0_known_answer():
    x = np.array([1.0, 2.0, 3.0, 4.0, 100.0])
    # sd = 43.60619, IQR (type 7) = 4 - 2 = 2 -> lo = 2 / 1.34; bw = 0.9 * lo * 5^-0.2
    assert abs(O.r_bw_nrd0(x) - 0.9 * (2.0 / 1.34) * 5 ** (-0.2)) < 1e-15
    assert O.r_bw_nrd0(np.array([3.0, 3.0, 3.0])) == 0.9 * 3.0 * 3 ** (-0.2)       # sd = IQR = 0 -> |x[1]|


def test_density_matches_the_exact_kernel_sum_and_integrates_to_one():
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.normal(0.0, 1.0, 3000), rng.normal(6.0, 0.5, 1000)])
    w = rng.random(x.size)
    w /= w.sum()
    for weights in (None, w):
        gx, gy = O.r_density(x, weights)
        bw = O.r_bw_nrd0(x)
        ww = np.full(x.size, 1.0 / x.size) if weights is None else weights
        exact = (ww[None, :] * np.exp(-0.5 * ((gx[:, None] - x[None, :]) / bw) ** 2)).sum(1) / (bw * np.sqrt(2 * np.pi))
        assert np.abs(gy - exact).max() < 5e-3 * exact.max()                      # linear binning onto 512 points
        assert abs(np.sum(0.5 * (gy[1:] + gy[:-1]) * np.diff(gx)) - 1.0) < 5e-3
        assert gx.size == 512 and gx[0] == x.min() - 3 * bw
    assert abs(O.r_density_mode(x) - 0.0) < 0.3 and abs(O.r_density_mode(x, np.where(x > 3, 50.0, 1.0)) - 6.0) < 0.2


def test_product_summary_table_matches_the_oracle_including_the_mode_row():
    """summary.py is pure torch (device sort, prefix-sum binning, torch.fft): check it on CPU tensors against the oracle."""
    import torch
    from abc_dls_b200 import summary as S
    rng = np.random.default_rng(9)
    n, P = 4000, 3
    vals = rng.normal(size=(n, P)) * np.array([1.0, 10.0, 0.01]) + np.array([0.0, 50.0, 1.0])
    w = rng.random(n)
    res = O.AbcOracleResult.__new__(O.AbcOracleResult)
    for method in ("rejection", "loclinear"):
        res.method, res.unadj_values, res.adj_values, res.weights = method, vals, vals, w
        want = O.summary_abc(res)
        ww = w if method == "loclinear" else np.ones(n)
        stats = np.stack([vals.min(0), vals.max(0), vals.sum(0), (ww[:, None] * vals).sum(0),
                          (ww[:, None] * vals * vals).sum(0), np.full(P, ww.sum())], axis=1)
        got = S.summary_table(torch.from_numpy(vals), torch.from_numpy(w), torch.from_numpy(stats),
                              weighted=(method == "loclinear"))
        np.testing.assert_allclose(np.asarray(got), want, rtol=1e-9, atol=1e-15)
        assert not np.isnan(np.asarray(got)).any()


def test_r_print_format_of_the_summary_table():
    """print.table / formatReal: one format per column, digits = 4 (what R shows for summary.abc)."""
    from abc_dls_b200 import summary as S
    assert O.r_format_real([0.1234567, 12.5, 1000.123]) == ["0.1235", "12.5000", "1000.1230"]
    assert O.r_format_real([1.234e-4, 150000.0]) == ["1.234e-04", "1.500e+05"]
    assert O.r_format_real([123456.0, 0.5]) == ["123456.0", "0.5"]
    assert O.r_format_real([-2.0, 3.14159]) == ["-2.000", "3.142"]
    assert O.r_format_real([0.0, 10.0]) == ["0", "10"]
    rng = np.random.default_rng(3)
    for _ in range(200):
        col = rng.normal(size=7) * 10.0 ** rng.integers(-6, 7)
        if rng.random() < 0.3:
            col = np.round(col, int(rng.integers(0, 3)))
        assert S.format_r_column(col) == O.r_format_real(col), col
    tab = S.SummaryTable(np.array([[1.0, 0.001], [2.5, 0.002], [3.25, 0.00345], [3.0, 0.004], [2.9, 0.0041],
                                   [4.0, 0.005], [10.0, 0.0123456]]), ["N_A", "mig"])
    txt = S.summary_text(tab, 100, True).splitlines()
    assert txt[0] == "Data:" and txt[-1].split() == ["Max.:", "10.00", "0.01235"]


# ---- leave-one-out tools: hand-checkable cases of the restatements -----------------------------------------------------
def test_cv_summaries_known_answers():
    cv = O.Cv4abcOracleResult(cvsamples=np.arange(3), tols=np.array([0.1]), true=np.array([[1.0], [2.0], [4.0]]),
                              estim={0.1: np.array([[1.5], [2.0], [3.0]])})
    # sum (est - true)^2 = .25 + 0 + 1 = 1.25; var(true) = 7/3 (n - 1 denominator); nval = 3
    assert abs(O.summary_cv4abc(cv)[0, 0] - 1.25 / (7.0 / 3.0 * 3)) < 1e-15
    cp = O.Cv4postprOracleResult(cvsamples=np.arange(4), tols=np.array([0.5]), true=np.array(["A", "A", "B", "B"]),
                                 models=["A", "B"],
                                 estim={0.5: np.array([[0.9, 0.1], [0.5, 0.5], [0.2, 0.8], [0.7, 0.3]])})
    s = O.summary_cv4postpr(cp)
    np.testing.assert_array_equal(s["conf.matrix"][0.5], [[2, 0], [1, 1]])       # the 0.5 / 0.5 tie goes to the first model
    np.testing.assert_allclose(s["probs"][0.5], [[0.7, 0.3], [0.45, 0.55]])
    g = O.GfitOracleResult(dist_obs=2.0, dist_sim=np.array([1.0, 3.0, 2.0, 5.0]))
    sg = O.summary_gfit(g)
    assert sg["pvalue"] == 0.5 and sg["dist.obs"] == 2.0                          # strictly greater than the observed
    np.testing.assert_allclose(sg["s.dist.sim"], [1.0, 1.75, 2.5, 2.75, 3.5, 5.0])   # Min 1stQu Median Mean 3rdQu Max


def test_leave_one_out_is_abc_on_the_table_without_the_row():
    """cv4abc / gfit restate a loop of abc() calls: one fit reproduced by hand (row removed, N - 1 rows in the MADs)."""
    rng = np.random.default_rng(21)
    par = rng.uniform(0, 10, size=(300, 2))
    ss = par + rng.normal(size=(300, 2))
    s = 123
    keep = np.arange(300) != s
    direct = O.abc(ss[s], par[keep], ss[keep], 0.1, "loclinear")
    assert direct.nacc == 30 and direct.idx.size == 30                     # ceiling(299 * 0.1)
    cv = O.cv4abc(par, ss, 1, 0.1, "loclinear", cvsamples=[s])
    np.testing.assert_array_equal(cv.estim[0.1][0], O.summary_abc(direct)[2])
    np.testing.assert_array_equal(cv.true[0], par[s])
    g = O.gfit(ss[5], ss, 1, 0.1, samples=[s])
    r = O.abc(ss[s], ss[keep, 0], ss[keep], 0.1, "rejection")
    assert g.dist_sim[0] == O.r_mean(r.dist[r.region])
    full = O.abc(ss[5], ss[:, 0], ss, 0.1, "rejection")
    assert g.dist_obs == O.r_mean(full.dist[full.region]) and full.dist[5] == 0.0   # the target is row 5 itself
