"""Generates tests/golden/*.npz: small seeded inputs with the outputs of the CPU oracle.

The reference itself cannot produce these (its ABC step is R `abc` 2.2.1 through rpy2; neither is installed, see
DESIGN.md §0), so the vectors pin the ORACLE: they guard it against regressions and give the GPU tests a fixed,
file-based target.  Re-run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import abc_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tables(n, d, P, seed):
    rng = np.random.default_rng(seed)
    lo, hi = rng.uniform(0, 10, P), rng.uniform(20, 1000, P)
    theta = (lo + (hi - lo) * rng.uniform(size=(n, P))).astype(np.float32).astype(np.float64)
    jm = np.arange(d) % P
    ss = (theta[:, jm] + 0.1 * (hi - lo)[jm] * rng.normal(size=(n, d))).astype(np.float32).astype(np.float64)
    return theta, ss, (lo + 0.37 * (hi - lo))[jm]


def main():
    # 1. joint loclinear (hcorr) and rejection
    theta, ss, tgt = tables(3000, 3, 2, 101)
    for method in ("loclinear", "rejection"):
        r = O.abc(tgt, theta, ss, 0.05, method)
        tab = O.summary_abc(r)
        np.savez_compressed(os.path.join(HERE, f"abc_{method}.npz"), target=tgt, param=theta, sumstat=ss, tol=0.05,
                            idx=r.idx, ds=r.ds, mad=r.mad, dist=r.dist, weights=r.weights, unadj=r.unadj_values,
                            adj=r.adj_values if r.adj_values is not None else np.zeros(0), summary=tab)
    # 2. ties: integer statistics, cap rule bites
    rng = np.random.default_rng(7)
    ssi = rng.integers(0, 6, size=(4000, 2)).astype(np.float64)
    par = rng.uniform(size=(4000, 1))
    r = O.abc(np.array([2.0, 3.0]), par, ssi, 0.02, "rejection")
    np.savez_compressed(os.path.join(HERE, "abc_ties.npz"), target=np.array([2.0, 3.0]), param=par, sumstat=ssi,
                        tol=0.02, idx=r.idx, ds=r.ds, mad=r.mad, n_le=int((r.dist <= r.ds).sum()))
    # 3. postpr, 5-decimal softmax
    n = 6000
    logits = rng.normal(size=(n, 3)) + 2.0 * (np.arange(3)[None, :] == (np.arange(n) % 3)[:, None])
    sm = np.round((np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32), 5).astype(np.float64)
    labels = np.array(["BNDX", "MNDX", "SNDX"])[np.arange(n) % 3]
    p = O.postpr([0.2, 0.5, 0.3], labels, sm, 0.05)
    np.savez_compressed(os.path.join(HERE, "postpr.npz"), target=np.array([0.2, 0.5, 0.3]), ids=(np.arange(n) % 3),
                        sumstat=sm, tol=0.05, idx=p.idx, counts=p.counts, ds=p.ds, mad=p.mad)
    # 4. one SMC round, per-column rejection (ABC.py:2709-2731)
    theta, ss, tgt = tables(5000, 4, 4, 303)
    hard_min, hard_max = theta.min(axis=0) - 1.0, theta.max(axis=0) + 1.0
    s = O.smc_round(tgt, theta, ss, 0.05, "rejection", theta, decrease=0.95, increase=0.01, hard_min=hard_min,
                    hard_max=hard_max)
    np.savez_compressed(os.path.join(HERE, "smc_round.npz"), target=tgt, param=theta, sumstat=ss, tol=0.05,
                        hard_min=hard_min, hard_max=hard_max, post_min=s["post_min"], post_max=s["post_max"],
                        new_min=s["min"], new_max=s["max"], decrease=s["decrease"], lines=s["lines"], lmrd=s["lmrd"])
    # 5. leave-one-out tools: per-column cv4abc (both methods), cv4postpr, gfit with fixed held-out rows
    theta, ss, tgt = tables(2500, 2, 2, 404)
    samp = np.array([0, 17, 1250, 2499, 801, 33])
    out = {"param": theta, "sumstat": ss, "target": tgt, "cvsamples": samp, "tols": np.array([0.02, 0.05])}
    for method in ("loclinear", "rejection"):
        for c in range(2):
            cv = O.cv4abc(theta[:, c], ss[:, c], samp.size, [0.02, 0.05], method, cvsamples=samp)
            out[f"{method}_col{c}_estim"] = np.stack([cv.estim[0.02], cv.estim[0.05]])
            out[f"{method}_col{c}_prederr"] = O.summary_cv4abc(cv)
    cvj = O.cv4abc(theta, ss, samp.size, [0.05], "rejection", cvsamples=samp)
    out["rejection_joint_estim"], out["rejection_joint_prederr"] = cvj.estim[0.05], O.summary_cv4abc(cvj)
    np.savez_compressed(os.path.join(HERE, "cv4abc.npz"), **out)
    n = 1500
    logits = rng.normal(size=(n, 3)) + 2.0 * (np.arange(3)[None, :] == (np.arange(n) % 3)[:, None])
    sm = np.round((np.exp(logits) / np.exp(logits).sum(1, keepdims=True)).astype(np.float32), 5).astype(np.float64)
    ids = np.arange(n) % 3
    sp = np.array([0, 300, 3, 1, 601, 1498, 2, 302, 1499])           # 3 per model, models in sorted order
    sp = np.concatenate([sp[ids[sp] == k] for k in range(3)])
    cp = O.cv4postpr(np.array(["BNDX", "MNDX", "SNDX"])[ids], sm, 3, [0.05], cvsamples=sp)
    sc = O.summary_cv4postpr(cp)
    gs = np.array([5, 77, 1400, 999])
    gf = O.gfit(np.array([0.2, 0.5, 0.3]), sm, gs.size, 0.05, samples=gs)
    sg = O.summary_gfit(gf)
    np.savez_compressed(os.path.join(HERE, "cv4postpr_gfit.npz"), ids=ids, sumstat=sm, cvsamples=sp, tol=0.05,
                        estim=cp.estim[0.05], conf=sc["conf.matrix"][0.05], probs=sc["probs"][0.05],
                        gfit_target=np.array([0.2, 0.5, 0.3]), gfit_samples=gs, dist_sim=gf.dist_sim, dist_obs=gf.dist_obs,
                        pvalue=sg["pvalue"], s_dist_sim=sg["s.dist.sim"])
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
