"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle — the reference's
R path cannot run here).  CPU: the oracle still reproduces them.  GPU: the CUDA path reproduces them."""
import os

import numpy as np
import pytest
import torch

from oracle import abc_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


@pytest.mark.parametrize("method", ["loclinear", "rejection"])
def test_oracle_reproduces_golden_abc(method):
    g = _load(f"abc_{method}.npz")
    r = O.abc(g["target"], g["param"], g["sumstat"], float(g["tol"]), method)
    np.testing.assert_array_equal(r.idx, g["idx"])
    np.testing.assert_array_equal(r.dist, g["dist"])
    assert r.ds == float(g["ds"])
    if method == "loclinear":
        np.testing.assert_allclose(r.adj_values, g["adj"], rtol=1e-12)
    np.testing.assert_allclose(O.summary_abc(r), g["summary"], rtol=1e-12, equal_nan=True)


def test_oracle_reproduces_golden_ties_postpr_smc():
    g = _load("abc_ties.npz")
    r = O.abc(g["target"], g["param"], g["sumstat"], float(g["tol"]), "rejection")
    np.testing.assert_array_equal(r.idx, g["idx"])
    assert int(g["n_le"]) > r.nacc
    g = _load("postpr.npz")
    p = O.postpr(g["target"], np.array(["BNDX", "MNDX", "SNDX"])[g["ids"]], g["sumstat"], float(g["tol"]))
    np.testing.assert_array_equal(p.counts, g["counts"])
    g = _load("smc_round.npz")
    s = O.smc_round(g["target"], g["param"], g["sumstat"], float(g["tol"]), "rejection", g["param"], 0.95, 0.01,
                    g["hard_min"], g["hard_max"])
    np.testing.assert_array_equal(s["min"], g["new_min"])
    np.testing.assert_array_equal(s["lines"], g["lines"])


def _tab(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["loclinear", "rejection"])
def test_cuda_reproduces_golden_abc(engine, method):
    g = _load(f"abc_{method}.npz")
    dev = engine.device
    r = engine.abc(torch.from_numpy(g["target"]).to(dev), _tab(g["param"], dev), _tab(g["sumstat"], dev),
                   float(g["tol"]), method, keep_dist=True)
    np.testing.assert_array_equal(r.idx.cpu().numpy(), g["idx"])
    np.testing.assert_array_equal(r.dist.cpu().numpy(), g["dist"])
    np.testing.assert_array_equal(r.mad.cpu().numpy(), g["mad"])
    assert r.ds == float(g["ds"])
    if method == "loclinear":
        np.testing.assert_array_equal(r.weights.cpu().numpy(), g["weights"])
        adj = r.adj_values.cpu().numpy()
        assert np.max(np.abs(adj - g["adj"]) / np.abs(g["adj"]).max(axis=0)) < 1e-9     # north_star tolerance
    mn, mx = r.minmax()
    np.testing.assert_allclose(mn.cpu().numpy(), g["summary"][0], rtol=1e-9)
    np.testing.assert_allclose(mx.cpu().numpy(), g["summary"][6], rtol=1e-9)


@pytest.mark.gpu
def test_cuda_reproduces_golden_ties_postpr_smc(engine):
    from abc_dls_b200 import smc
    dev = engine.device
    g = _load("abc_ties.npz")
    r = engine.abc(torch.from_numpy(g["target"]).to(dev), _tab(g["param"], dev), _tab(g["sumstat"], dev),
                   float(g["tol"]), "rejection")
    np.testing.assert_array_equal(r.idx.cpu().numpy(), g["idx"])
    g = _load("postpr.npz")
    p = engine.postpr(torch.from_numpy(g["target"]).to(dev), torch.from_numpy(g["ids"]).to(torch.int32),
                      ["BNDX", "MNDX", "SNDX"], _tab(g["sumstat"], dev), float(g["tol"]))
    np.testing.assert_array_equal(p.counts.cpu().numpy(), g["counts"])
    np.testing.assert_array_equal(p.idx.cpu().numpy(), g["idx"])
    # one SMC after-train round on the device: per-column rejection -> min/max -> range rules -> box filter
    g = _load("smc_round.npz")
    par, ss = _tab(g["param"], dev), _tab(g["sumstat"], dev)
    res = engine.abc_columns(torch.from_numpy(g["target"]).to(dev), par, ss, float(g["tol"]), "rejection")
    pmin = np.array([float(x.minmax()[0][0]) for x in res])
    pmax = np.array([float(x.minmax()[1][0]) for x in res])
    np.testing.assert_array_equal(pmin, g["post_min"])
    np.testing.assert_array_equal(pmax, g["post_max"])
    omin, omax = engine.oldrange(par)
    rng = smc.updating_newrange(pmin, pmax, omin.cpu().numpy(), omax.cpu().numpy(), 0.95)
    rng = smc.noise_injection_update(rng, omin.cpu().numpy(), omax.cpu().numpy(), 0.01, g["hard_min"], g["hard_max"], 0.95)
    np.testing.assert_array_equal(rng.min, g["new_min"])
    np.testing.assert_array_equal(rng.max, g["new_max"])
    np.testing.assert_array_equal(rng.decrease, g["decrease"])
    assert smc.lmrd_calculation(rng, g["hard_min"], g["hard_max"]) == float(g["lmrd"])
    idx = engine.box_filter(par, torch.from_numpy(rng.min), torch.from_numpy(rng.max))
    np.testing.assert_array_equal(idx.cpu().numpy() + 2, g["lines"])


# ---- leave-one-out tools -------------------------------------------------------------------------------------------------
def test_oracle_reproduces_golden_cv():
    g = _load("cv4abc.npz")
    for method in ("loclinear", "rejection"):
        for c in range(2):
            cv = O.cv4abc(g["param"][:, c], g["sumstat"][:, c], g["cvsamples"].size, g["tols"], method,
                          cvsamples=g["cvsamples"])
            np.testing.assert_allclose(np.stack([cv.estim[0.02], cv.estim[0.05]]), g[f"{method}_col{c}_estim"], rtol=1e-12)
            np.testing.assert_allclose(O.summary_cv4abc(cv), g[f"{method}_col{c}_prederr"], rtol=1e-10)
    g = _load("cv4postpr_gfit.npz")
    labels = np.array(["BNDX", "MNDX", "SNDX"])[g["ids"]]
    cp = O.cv4postpr(labels, g["sumstat"], 3, [float(g["tol"])], cvsamples=g["cvsamples"])
    np.testing.assert_array_equal(cp.estim[0.05], g["estim"])
    np.testing.assert_array_equal(O.summary_cv4postpr(cp)["conf.matrix"][0.05], g["conf"])
    gf = O.gfit(g["gfit_target"], g["sumstat"], g["gfit_samples"].size, float(g["tol"]), samples=g["gfit_samples"])
    np.testing.assert_allclose(gf.dist_sim, g["dist_sim"], rtol=1e-13)
    assert O.summary_gfit(gf)["pvalue"] == float(g["pvalue"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_cv(engine):
    from abc_dls_b200 import cv as CV
    g = _load("cv4abc.npz")
    dev = engine.device
    par, ss = _tab(g["param"], dev), _tab(g["sumstat"], dev)
    scale = np.abs(g["param"]).max(axis=0)
    for method in ("loclinear", "rejection"):
        for c in range(2):          # one thread block per fit (csrc/small.cu)
            cv = CV.cv4abc(engine, par[c:c + 1], ss[c:c + 1], g["cvsamples"].size, g["tols"], method,
                           cvsamples=g["cvsamples"])
            got = np.stack([cv.estim[0.02], cv.estim[0.05]])
            assert np.max(np.abs(got - g[f"{method}_col{c}_estim"])) <= 1e-9 * scale[c]
            np.testing.assert_allclose(cv.summary(print_=False), g[f"{method}_col{c}_prederr"], rtol=1e-6)
    cvj = CV.cv4abc(engine, par, ss, g["cvsamples"].size, [0.05], "rejection", cvsamples=g["cvsamples"])   # streaming
    np.testing.assert_array_equal(cvj.estim[0.05], g["rejection_joint_estim"])
    g = _load("cv4postpr_gfit.npz")
    ids = torch.from_numpy(g["ids"].astype(np.int32)).to(dev)
    ssm = _tab(g["sumstat"], dev)
    cp = CV.cv4postpr(engine, ids, ["BNDX", "MNDX", "SNDX"], ssm, 3, [float(g["tol"])], cvsamples=g["cvsamples"])
    np.testing.assert_array_equal(cp.estim[0.05], g["estim"])
    s = cp.summary(print_=False)
    np.testing.assert_array_equal(s["conf.matrix"][0.05], g["conf"])
    np.testing.assert_allclose(s["probs"][0.05], g["probs"], rtol=1e-12)
    gf = CV.gfit(engine, torch.from_numpy(g["gfit_target"]).to(dev), ssm, g["gfit_samples"].size, float(g["tol"]),
                 samples=g["gfit_samples"])
    np.testing.assert_allclose(gf.dist_sim, g["dist_sim"], rtol=1e-12)
    assert abs(gf.dist_obs - float(g["dist_obs"])) <= 1e-12 * float(g["dist_obs"])
    assert gf.summary(print_=False)["pvalue"] == float(g["pvalue"])
