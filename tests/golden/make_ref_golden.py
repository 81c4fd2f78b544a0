"""Generates the fixtures that pin this repo against THE REFERENCE ITSELF, run in the build container:

    python tests/golden/make_ref_golden.py          (needs /root/reference; writes tests/golden/ref_*.{npz,json})

1. ``ref_smc_rules.npz`` - outputs of the reference's own, unmodified SMC range functions
   (``ABC_DLS_SMC.updating_newrange`` ABC.py:2921-2946, ``noise_injection_update`` 2881-2919, ``lmrd_calculation``
   3016-3028, ``narrowing_params`` 2974-2990; ``ABC_DLS_SMC_Snakemake.median_newrange`` /
   ``update_newrange_using_oldrange_hardrange`` / ``lmrd4mcsv`` / ``remove_repeated_params`` SFS/Class.py:706-863) on
   seeded inputs incl. the edge cases of those rules.  ``abc_dls_b200.smc`` must reproduce them bit for bit.
2. ``ref_callsites.json`` - what the reference's unmodified call sites print / return when they run over
   ``abc_dls_b200.rabc`` + ``rshim`` (bound exactly as INTEGRATION.md says; the arithmetic underneath is the numpy
   stage double, there is no GPU here): ``abc_params`` (ABC.py:1929-1973), ``abc_params_min_max`` (2788-2861),
   ``model_selection`` (901-921), ``plot_power_of_ss`` (863-899), ``goodness_fit`` (946-965), ``plot_param_cv_error``
   (1877-1927).  The ``-m gpu`` tests replay the same abc()/postpr()/cv calls on the CUDA engine from the stored inputs
   and must print the same text.

The reference's arithmetic for abc()/postpr() is R's (absent here), so (2) pins the BOUNDARY (argument conventions,
what is read from a result, text layout), not R's numbers; (1) pins arithmetic, because those rules are the reference's
own Python.
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refload  # noqa: E402


def smc_rule_cases():
    """Seeded inputs of the range rules: random boxes + the rows that trigger every branch."""
    cases = []
    for seed in range(24):
        rng = np.random.default_rng(1000 + seed)
        P = int(rng.integers(3, 25))
        omin = np.round(rng.uniform(0, 10, P) * 10.0 ** rng.integers(-2, 4, P), 4)
        omax = omin + np.round(rng.uniform(1, 100, P) * 10.0 ** rng.integers(-2, 4, P), 4)
        frac = rng.uniform(0.3, 1.0, P)
        pmin = omin + (omax - omin) * (1 - frac) / 2
        pmax = pmin + (omax - omin) * frac
        k = rng.permutation(P)
        pmin[k[0]], pmax[k[0]] = omin[k[0]], omax[k[0]]              # decrease == 1 exactly
        pmin[k[1]] = pmax[k[1]] = omin[k[1]] + 0.5 * (omax[k[1]] - omin[k[1]])     # decrease == 0 -> whole row reverts
        if P > 3:
            j = k[2]
            omin[j], omax[j], pmin[j], pmax[j] = 5.0, 5.001, 5.0001, 5.0002      # collapses under round(3)
        hmin = omin - np.abs(omin) * rng.uniform(0, 0.2, P)
        hmax = omax + np.abs(omax) * rng.uniform(0, 0.2, P)
        if seed % 3 == 0:                                             # hard range tighter than the widened box
            hmin, hmax = omin.copy(), omax.copy()
        cases.append(dict(P=P, omin=omin, omax=omax, pmin=pmin, pmax=pmax, hmin=hmin, hmax=hmax,
                          decrease=[0.95, 0.8, 0.5][seed % 3], increase=[0.0, 0.01, 0.05, 0.2][seed % 4],
                          hard=seed % 2 == 0))
    return cases


def make_smc_rules(ABC, SC, out):
    S, K = ABC.ABC_DLS_SMC, SC.ABC_DLS_SMC_Snakemake
    blob = {}
    for i, c in enumerate(smc_rule_cases()):
        names = [f"p{j}" for j in range(c["P"])]
        new = pd.DataFrame({"min": c["pmin"], "max": c["pmax"]}, index=names)
        old = pd.DataFrame({"min": c["omin"], "max": c["omax"]}, index=names)
        hard = pd.DataFrame({"min": c["hmin"], "max": c["hmax"]}, index=names)
        upd = S.updating_newrange(newrange=new.copy(), oldrange=old, decrease=c["decrease"])
        for k_ in ("omin", "omax", "pmin", "pmax", "hmin", "hmax"):
            blob[f"c{i}_{k_}"] = c[k_]
        blob[f"c{i}_cfg"] = np.array([c["decrease"], c["increase"], float(c["hard"])])
        blob[f"c{i}_upd"] = upd[["min", "max", "decrease"]].to_numpy()
        if c["increase"] > 0:
            inj = S.noise_injection_update(newrange=upd.copy(), oldrange=old, increase=c["increase"],
                                           hardrange=hard if c["hard"] else pd.DataFrame(), decrease=c["decrease"])
            blob[f"c{i}_inj"] = inj[["min", "max", "decrease"]].to_numpy()
            blob[f"c{i}_lmrd"] = np.array([S.lmrd_calculation(newrange=inj, hardrange=hard)])
        else:
            blob[f"c{i}_lmrd"] = np.array([S.lmrd_calculation(newrange=upd, hardrange=hard)])
    blob["n_cases"] = np.array([len(smc_rule_cases())])

    # narrowing_params on a table with values exactly on the box faces (Series.between is inclusive)
    rng = np.random.default_rng(7)
    par = np.round(rng.uniform(0, 1, size=(5000, 4)), 3)
    lo, hi = np.array([0.1, 0.25, 0.0, 0.5]), np.array([0.9, 0.75, 1.0, 0.999])
    blob["narrow_par"], blob["narrow_lo"], blob["narrow_hi"] = par, lo, hi
    blob["narrow_lines"] = np.asarray(S.narrowing_params(params=pd.DataFrame(par), parmin=pd.Series(lo),
                                                         parmax=pd.Series(hi)), dtype=np.int64)

    # Snakemake glue on real files (header-less csv: name,min,max,decrease)
    with tempfile.TemporaryDirectory() as td:
        c = smc_rule_cases()[5]
        names = [f"p{j}" for j in range(c["P"])]
        files = []
        rng = np.random.default_rng(99)
        for r in range(4):                                            # even count: median averages the middle two
            f = os.path.join(td, f"Newrange_{r}.csv")
            jit = rng.uniform(-0.02, 0.02, (2, c["P"])) * (c["omax"] - c["omin"])
            pd.DataFrame({"min": np.round(c["pmin"] + jit[0], 3), "max": np.round(c["pmax"] + jit[1], 3),
                          "decrease": 0.5}, index=names).to_csv(f, header=False)
            files.append(f)
        blob["snk_files"] = np.stack([pd.read_csv(f, header=None).iloc[:, 1:3].to_numpy() for f in files])
        blob["snk_names"] = np.array(names)
        med = K.median_newrange(files)
        blob["snk_median"] = med[["min", "max"]].to_numpy()
        old = pd.DataFrame({"min": c["omin"], "max": c["omax"]}, index=names)
        hf = os.path.join(td, "hardrange.csv")
        # the hard-range file lists the parameters in ANOTHER order: pandas aligns on the names
        order = rng.permutation(c["P"])
        pd.DataFrame({"min": c["hmin"][order], "max": c["hmax"][order]},
                     index=[names[j] for j in order]).to_csv(hf, header=False)
        blob["snk_hard_order"] = order
        upd = K.update_newrange_using_oldrange_hardrange(newrange=med.copy(), oldrange=old, decrease=0.9,
                                                         increase=0.02, hardrange_file=hf)
        blob["snk_case"] = np.array([5, 0.9, 0.02])
        blob["snk_updated"] = upd[["min", "max", "decrease"]].to_numpy()
        nf = os.path.join(td, "Newrange.csv")
        upd.to_csv(nf, header=False)
        blob["snk_newrange_csv"] = np.array(open(nf).read())
        blob["snk_lmrd"] = np.array([K.lmrd4mcsv(hardrange_file=hf, newrange_file=nf)])
        # remove_repeated_params (reference behaviour incl. its line shift, see abc_dls_b200/smc.py)
        allcsv = os.path.join(td, "All.csv")
        vals = rng.integers(0, 4, size=(60, 2))
        lines = [f"{a},{b},row{i}" for i, (a, b) in enumerate(vals)]
        open(allcsv, "w").write("\n".join(lines) + "\n")
        K.remove_repeated_params(inputfile=allcsv, paramsnumbers=2, outputfile=os.path.join(td, "dedup.csv"))
        blob["dedup_in"] = np.array("\n".join(lines) + "\n")
        blob["dedup_out"] = np.array(open(os.path.join(td, "dedup.csv")).read())
    np.savez_compressed(out, **blob)
    print("wrote", out)


class Recorder:
    """Sits where rpy2's package handle sat (ABC.abc): forwards every attribute to rabc and keeps what the calls
    returned, so that the generator can store e.g. the held-out rows a cv4abc() drew."""

    def __init__(self, mod):
        self._mod, self.log = mod, []

    def __getattr__(self, name):
        fn = getattr(self._mod, name)
        if not callable(fn):
            return fn

        def wrapped(*a, **k):
            res = fn(*a, **k)
            self.log.append((name, res))
            return res
        return wrapped


def callsite_inputs():
    """float32-representable tables like the ones ABC-DLS hands over (Keras predictions widened to double)."""
    rng = np.random.default_rng(20240919)
    N, names = 3000, ["N_A", "T_split", "mig"]
    lo, hi = np.array([5000.0, 10000.0, 0.0]), np.array([20000.0, 120000.0, 0.02])
    par = (lo + (hi - lo) * rng.random((N, 3))).astype(np.float32).astype(np.float64)
    pred = (par + 0.1 * (hi - lo) * rng.standard_normal((N, 3))).astype(np.float32)
    target = (lo + 0.37 * (hi - lo))[None, :]
    K = 3
    mid = np.arange(N) % K
    logits = 2.0 * np.eye(K)[mid] + rng.standard_normal((N, K))
    sm = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    return dict(names=names, par=par, pred=pred, target=target, models=["BNDX", "MNDX", "SNDX"], mid=mid,
                ss_post=np.round(sm, 5), target_post=np.array([[0.2, 0.5, 0.3]]))


def capture(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ret = fn(*a, **k)
    return ret, buf.getvalue()


def make_callsites(ABC, out_json, out_npz):
    from abc_dls_b200 import rabc
    from abc_dls_b200.engine import AbcEngine
    from cpu_stages import CpuStages
    rabc.set_engine(AbcEngine(stages=CpuStages()))
    rabc.default_seed = 4242
    rec = Recorder(rabc)
    ABC.abc = rec
    inp = callsite_inputs()
    np.savez_compressed(out_npz, par=inp["par"], pred=inp["pred"], target=inp["target"], mid=inp["mid"],
                        ss_post=inp["ss_post"], target_post=inp["target_post"])
    names = inp["names"]
    param = pd.DataFrame(inp["par"], columns=names)
    pred = pd.DataFrame(inp["pred"], columns=names)
    target = pd.DataFrame(inp["target"], columns=names)
    index = pd.Series(np.array(inp["models"])[inp["mid"]])
    ss_post = pd.DataFrame(inp["ss_post"])
    target_post = pd.DataFrame(inp["target_post"])
    tol = 0.05
    g = {"tol": tol, "names": names, "models": inp["models"], "seed": rabc.default_seed}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)          # the reference writes its sink files into the working directory
        try:
            for method in ("rejection", "loclinear"):
                _, g[f"abc_params_{method}"] = capture(ABC.ABC_DLS_Params.abc_params, target=target, param=param,
                                                       ss=pred, tol=tol, method=method, name="post.pdf")
                (mn, mx), _ = capture(ABC.ABC_DLS_SMC.abc_params_min_max, target=target, param=param, ss=pred,
                                      tol=tol, method=method)
                g[f"minmax_{method}"] = [np.asarray(mn, dtype=np.float64).tolist(),
                                         np.asarray(mx, dtype=np.float64).tolist()]
            _, g["model_selection_rejection"] = capture(ABC.ABC_DLS_Classification.model_selection, target=target_post,
                                                        ss=ss_post, index=index, tol=tol, method="rejection")
            rec.log.clear()
            _, g["plot_power_of_ss_rejection"] = capture(ABC.ABC_DLS_Classification.plot_power_of_ss, ss=ss_post,
                                                         index=index, tol=tol, repeats=4, method="rejection")
            g["cv4postpr_cvsamples"] = np.asarray(rec.log[-1][1].cvsamples).tolist()
            rec.log.clear()
            _, g["goodness_fit"] = capture(ABC.ABC_DLS_Classification.goodness_fit, target=target_post,
                                           ss=ss_post.iloc[inp["mid"] == 1], name="MNDX", tol=tol, repeats=5)
            g["gfit_dist_sim"] = np.asarray(rec.log[-1][1].dist_sim).tolist()
            for method in ("rejection", "loclinear"):
                rec.log.clear()
                ABC.robjects.r["sink"]()     # the reference leaves its last sink open (ABC.py:1917-1923); R would too
                _, g[f"plot_param_cv_error_{method}"] = capture(ABC.ABC_DLS_Params.plot_param_cv_error, param=param,
                                                                ss=pred, tol=tol, repeats=5, method=method,
                                                                name="cv.pdf")
                g[f"cv4abc_cvsamples_{method}"] = [np.asarray(r.cvsamples).tolist() for n_, r in rec.log
                                                   if n_ == "cv4abc"]
            ABC.robjects.r["sink"]()
        finally:
            os.chdir(cwd)
            ABC.abc = rabc
            rabc.default_seed = None
            rabc.set_engine(None)
    with open(out_json, "w") as fh:
        json.dump(g, fh, indent=1)
    print("wrote", out_json, out_npz)


if __name__ == "__main__":
    ABC, SC = refload.load()
    make_smc_rules(ABC, SC, os.path.join(HERE, "ref_smc_rules.npz"))
    make_callsites(ABC, os.path.join(HERE, "ref_callsites.json"), os.path.join(HERE, "ref_callsites_inputs.npz"))
