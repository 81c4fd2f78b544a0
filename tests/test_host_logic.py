"""Host-side logic on CPU: SMC range rules, summary table semantics, and the engine's sharded sequencing of
stages and collectives under gloo with world_size 2 (numpy test double for the CUDA stages, tests/cpu_stages.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import abc_oracle as O  # noqa: E402
from abc_dls_b200 import smc, summary, synth  # noqa: E402


def test_smc_rules_match_oracle():
    rng = np.random.default_rng(21)
    P = 24
    omin = rng.uniform(0, 10, P)
    omax = omin + rng.uniform(1, 100, P)
    frac = rng.uniform(0.3, 1.0, P)
    pmin = omin + (omax - omin) * (1 - frac) / 2
    pmax = pmin + (omax - omin) * frac
    pmin[3], pmax[3] = omin[3], omax[3]
    pmin[5] = pmax[5] = omin[5] + 1.0
    pmin[7], pmax[7], omin[7], omax[7] = 5.0001, 5.0002, 5.0, 5.001
    hmin, hmax = omin - 1.0, omax + 0.5
    r = smc.updating_newrange(pmin, pmax, omin, omax, 0.95)
    o = O.updating_newrange(pmin, pmax, omin, omax, 0.95)
    for a, b in zip((r.min, r.max, r.decrease), o):
        np.testing.assert_array_equal(a, b)
    r2 = smc.noise_injection_update(r, omin, omax, 0.01, hmin, hmax, 0.95)
    o2 = O.noise_injection_update(*o, omin, omax, 0.01, hmin, hmax, 0.95)
    for a, b in zip((r2.min, r2.max, r2.decrease), o2):
        np.testing.assert_array_equal(a, b)
    assert smc.lmrd_calculation(r2, hmin, hmax) == O.lmrd_calculation(o2[0], o2[1], hmin, hmax)


def test_summary_table_indexing_follows_r():
    tab = summary.SummaryTable(np.arange(14.0).reshape(7, 2), ["a", "b"])
    assert tab[0] == 0.0 and tab[6] == 12.0            # column-major flat index: Min / Max of parameter 1
    assert tab[7] == 1.0 and tab[13] == 13.0           # ... of parameter 2
    assert np.array(tab).shape == (7, 2)
    np.testing.assert_array_equal(np.array(tab)[0], [0.0, 1.0])     # ABC.py:2854 reads row 0 / row -1
    txt = summary.summary_text(tab, 10, True)
    assert txt.startswith("Data:") and "Weights:" in txt and "Min.:" in txt


def _worker(rank, world, port, N, d, P, tol, method, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from abc_dls_b200.engine import AbcEngine
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from cpu_stages import CpuStages
    eng = AbcEngine(stages=CpuStages())
    per = (N + world - 1) // world
    r0, r1 = rank * per, min(N, (rank + 1) * per)
    par, ss, tgt = synth.abc_tables(N, d, P, "cpu")
    res = eng.abc(tgt, par[:, r0:r1].contiguous(), ss[:, r0:r1].contiguous(), tol, method)
    mn, mx = res.minmax()
    q.put((rank, res.idx.numpy(), res.values.contiguous().numpy(), res.ds, res.mad.numpy(), mn.numpy(), mx.numpy(),
           eng.comm.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("method,tol", [("rejection", 0.03), ("loclinear", 0.05)])
def test_sharded_sequencing_world2_gloo(method, tol):
    N, d, P, world = 4001, 3, 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + (1 if method == "rejection" else 2)
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, d, P, tol, method, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    par, ss, tgt = synth.abc_tables(N, d, P, "cpu")
    o = O.abc(tgt.numpy(), par.numpy().T, ss.numpy().T, tol, method)
    idx = np.concatenate([t[1] for t in out])
    vals = np.concatenate([t[2] for t in out])
    np.testing.assert_array_equal(idx, o.idx)                # bit-exact accepted set, global row order across ranks
    assert out[0][3] == o.ds and out[1][3] == o.ds
    np.testing.assert_array_equal(out[0][4], o.mad)
    ref = o.adj_values if method == "loclinear" else o.unadj_values
    assert np.max(np.abs(vals - ref) / np.abs(ref).max(axis=0)) < 1e-9
    for t in out:                                            # global summary on every rank
        np.testing.assert_allclose(t[5], ref.min(axis=0), rtol=1e-9)
        np.testing.assert_allclose(t[6], ref.max(axis=0), rtol=1e-9)
        assert t[7] >= 12                                    # the radix passes alone need 18 all-reduces


def test_snakemake_range_files_roundtrip(tmp_path):
    """§8f-3: median of repeated Newrange.csv files -> update rules -> Newrange.csv, against the reference's pandas
    formulation (src/SFS/Class.py:706-796, 850-863) restated inline."""
    import pandas
    rng = np.random.default_rng(8)
    names = ["N_A", "N_AF", "T_B", "m_AF_B"]
    omin, omax = np.array([100.0, 1000.0, 0.0, 0.0]), np.array([5e4, 1e5, 1.0, 0.02])
    files = []
    for k in range(4):                       # even count: medians average the two middle files
        lo = omin + (omax - omin) * rng.uniform(0.0, 0.3, 4)
        hi = omax - (omax - omin) * rng.uniform(0.0, 0.3, 4)
        if k == 0:
            lo[2], hi[2] = omin[2], omax[2]
        f = tmp_path / f"Newrange_{k}.csv"
        smc.RangeTable(lo, hi, (hi - lo) / (omax - omin)).to_csv(str(f), names)
        files.append(str(f))
    hard = tmp_path / "hardrange.csv"
    hmin, hmax = omin + 1.0, omax * 0.99
    with open(hard, "w") as fh:
        for n, a, b in reversed(list(zip(names, hmin, hmax))):          # other row order: aligned by name
            fh.write(f"{n},{a},{b}\n")
    out = tmp_path / "Newrange.csv"
    got_names, got = smc.multiple_newrange2updatednewrange(files, omin, omax, 0.95, 0.02, str(hard), str(out))

    alldata = [pandas.read_csv(f, header=None).iloc[:, :3] for f in files]
    mn = pandas.DataFrame([d.iloc[:, 1] for d in alldata]).median().to_numpy()
    mx = pandas.DataFrame([d.iloc[:, 2] for d in alldata]).median().to_numpy()
    w = O.updating_newrange(mn, mx, omin, omax, 0.95)
    w = O.noise_injection_update(w[0], w[1], w[2], omin, omax, 0.02, hmin, hmax, 0.95)
    assert got_names == names
    np.testing.assert_array_equal(got.min, w[0]); np.testing.assert_array_equal(got.max, w[1])
    np.testing.assert_array_equal(got.decrease, w[2])
    back = pandas.read_csv(out, header=None, index_col=0)
    np.testing.assert_array_equal(back.iloc[:, 0].to_numpy(), w[0]); np.testing.assert_array_equal(back.iloc[:, 2].to_numpy(), w[2])
    assert smc.lmrd4mcsv(str(hard), str(out)) == O.lmrd_calculation(w[0], w[1], hmin, hmax)
    np.testing.assert_array_equal(smc.csv_linenumbers(np.array([0, 5])), [2, 7])
    np.testing.assert_array_equal(smc.csv_linenumbers(np.array([0, 5]), header=False), [1, 6])


def test_row_major_param_sequencing_on_the_stage_double():
    """engine.py routes a (P, n) view with strides (1, P) through gather(param=None) + gather_param_rows."""
    from abc_dls_b200.engine import AbcEngine
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from cpu_stages import CpuStages
    eng = AbcEngine(stages=CpuStages())
    par, ss, tgt = synth.abc_tables(3000, 3, 4, "cpu")
    rm = par.t().contiguous()
    a = eng.abc(tgt, par, ss, 0.05, "loclinear")
    b = eng.abc(tgt, rm.t(), ss, 0.05, "loclinear")
    assert torch.equal(a.idx, b.idx) and torch.equal(a.adj_values, b.adj_values) and torch.equal(a.stats, b.stats)
    o = O.abc(tgt.numpy(), par.numpy().T, ss.numpy().T, 0.05, "loclinear")
    np.testing.assert_array_equal(b.idx.numpy(), o.idx)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the contract
    keys, the same metric / unit / config as the CUDA arm, timed on a bounded sample."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-rows", "20000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "sims/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["config"]["rows"] == 100000000


def _worker_postpr_smc(rank, world, port, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from abc_dls_b200.engine import AbcEngine
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from cpu_stages import CpuStages
    eng = AbcEngine(stages=CpuStages())
    per = (N + world - 1) // world
    r0, r1 = rank * per, min(N, (rank + 1) * per)
    ids, ss, tgt = synth.postpr_tables(N, "cpu")
    ss = torch.round(ss, decimals=2)                      # coarse grid: many rows tie exactly at ds
    pp = eng.postpr(tgt, ids[r0:r1].contiguous(), ["BNDX", "MNDX", "SNDX"], ss[:, r0:r1].contiguous(), 0.05)
    par, _, _ = synth.abc_tables(N, 1, 5, "cpu")
    mn, mx = eng.oldrange(par[:, r0:r1].contiguous())
    lo, hi = mn + 0.2 * (mx - mn), mx - 0.1 * (mx - mn)
    keep = eng.box_filter(par[:, r0:r1].contiguous(), lo, hi)
    # string model labels: the level table is the union over ranks (rank 0 never sees "SNDX", rank 1 never "BNDX")
    from abc_dls_b200 import rabc
    codes, levels = rabc._model_ids(np.array([["BNDX", "MNDX"], ["SNDX", "MNDX"]][rank] * 3), eng)
    assert levels == ["BNDX", "MNDX", "SNDX"] and codes.tolist() == [[0, 1], [2, 1]][rank] * 3
    _, klev = rabc._model_ids(torch.tensor([0, 1 + rank], dtype=torch.int32), eng)
    assert klev == ["0", "1", "2"]                       # integer ids: the largest id over ALL ranks sizes the table
    q.put((rank, pp.counts.numpy(), pp.idx.numpy(), pp.ds, mn.numpy(), mx.numpy(), keep.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_postpr_and_smc_passes_sharded_world2_gloo():
    """postpr (5-decimal softmax: exact ties across the shard boundary, so the cross-rank cap rule decides), oldrange and
    the box filter over two ranks: global row numbers, global extremes."""
    N, world = 6001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_worker_postpr_smc, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids, ss, tgt = synth.postpr_tables(N, "cpu")
    ss = torch.round(ss, decimals=2)
    o = O.postpr(tgt.numpy(), np.array(["BNDX", "MNDX", "SNDX"])[ids.numpy()], ss.numpy().T, 0.05)
    for t in out:
        np.testing.assert_array_equal(t[1], o.counts)                       # global counts on every rank
        assert t[3] == o.ds
    np.testing.assert_array_equal(np.concatenate([t[2] for t in out]), o.idx)
    assert (o.dist <= o.ds).sum() > o.nacc                                   # ties at ds: the cap rule was exercised
    par, _, _ = synth.abc_tables(N, 1, 5, "cpu")
    p = par.numpy()
    for t in out:
        np.testing.assert_array_equal(t[4], p.min(axis=1))
        np.testing.assert_array_equal(t[5], p.max(axis=1))
    lo, hi = p.min(axis=1) + 0.2 * (p.max(axis=1) - p.min(axis=1)), p.max(axis=1) - 0.1 * (p.max(axis=1) - p.min(axis=1))
    np.testing.assert_array_equal(np.concatenate([t[6] for t in out]) + 2, O.narrowing_params(p.T, lo, hi))


def test_postpr_summary_text_uses_r_formats():
    from abc_dls_b200 import rabc
    from abc_dls_b200.engine import PostprResult
    counts = torch.tensor([40, 187, 74])
    res = PostprResult(models=["BNDX", "MNDX", "SNDX"], counts=counts, pred=counts.double() / counts.sum(), nacc=301,
                       n_global=6001, ds=0.67, mad=torch.ones(3), idx=torch.arange(301), ss=torch.zeros(301, 3),
                       values=torch.zeros(301, dtype=torch.int32))
    fit = rabc.PostprFit(res)
    txt = fit.summary_text().splitlines()
    assert txt[0] == "Data:" and txt[1] == " postpr.out$values (301 posterior samples)"
    i = txt.index("Proportion of accepted simulations (rejection):")
    assert txt[i + 1].split() == ["BNDX", "MNDX", "SNDX"] and txt[i + 2].split() == ["0.1329", "0.6213", "0.2458"]
    j = txt.index("Bayes factors:")
    assert txt[j + 2].split() == ["BNDX", "1.000", "0.2139", "0.5405"]       # one format per column, 4 significant digits
    assert txt[j + 3].split() == ["MNDX", "4.675", "1.0000", "2.5270"]
    s = fit.summary(print_=False)
    assert abs(s["Prob"]["MNDX"] - 187 / 301) < 1e-15 and s["BayesF"].shape == (3, 3)


def test_reference_summary_flows_through_the_shim(tmp_path, capsys):
    """The three ways ABC-DLS reads an abc() result (ABC.py:844-861 r_summary, 2821-2840 per-parameter min / max,
    2842-2861 all parameters), driven exactly like the reference does - sink to a temp file, summary(), sink() - over
    the rpy2 stand-ins, with the engine's sequencing running on the numpy stage double."""
    from abc_dls_b200 import rabc
    from abc_dls_b200.rshim import r
    from abc_dls_b200.engine import AbcEngine
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from cpu_stages import CpuStages
    eng = AbcEngine(stages=CpuStages())
    par, ss, tgt = synth.abc_tables(4000, 2, 2, "cpu")
    o = O.abc(tgt.numpy(), par.numpy().T, ss.numpy().T, 0.05, "loclinear")
    want = O.summary_abc(o)
    fit = rabc.AbcFit(eng.abc(tgt, par, ss, 0.05, "loclinear"), eng, ["N_A", "T_B"], ["s1", "s2"])

    temp = str(tmp_path / "summary.txt")
    r.options(width=10000)
    r["sink"](temp)
    r["summary"](fit)
    r["sink"]()
    lines = open(temp).read().splitlines()
    k = [i for i, l in enumerate(lines) if l.startswith("Data:")][0]
    assert lines[k + 1].strip() == "abc.out$adj.values (200 posterior samples)" and lines[k + 2] == "Weights:"
    body = [l for l in lines if l.startswith("Min.:") or l.startswith("Max.:")]
    assert len(body) == 2 and len(body[0].split()) == 3

    r["sink"](temp)
    single_min = r["summary"](fit)[0]
    r["sink"]()
    r["sink"](temp)
    single_max = r["summary"](fit)[6]
    r["sink"]()
    assert abs(single_min - want[0, 0]) <= 1e-9 * abs(want[0, 0]) and abs(single_max - want[6, 0]) <= 1e-9 * abs(want[6, 0])

    import pandas
    r["sink"](temp)
    min_df = pandas.DataFrame(np.array(r["summary"](fit))).iloc[0, :]
    r["sink"]()
    r["sink"](temp)
    max_df = pandas.DataFrame(np.array(r["summary"](fit))).iloc[-1, :]
    r["sink"]()
    np.testing.assert_allclose(min_df.to_numpy(), want[0], rtol=1e-9)
    np.testing.assert_allclose(max_df.to_numpy(), want[6], rtol=1e-9)
    np.testing.assert_allclose(np.array(fit.summary(print_=False)), want, rtol=1e-9)
    assert capsys.readouterr().out == ""                   # everything summary() printed went to the sink


def test_unsupported_methods_are_delegated_to_r_or_stop_with_the_reason():
    """`rabc` replaces the rpy2 handle wholesale, and 'neuralnet' is the default --method of two
    Run_ParamsEstimation sub-commands: such calls go to R's abc package through rpy2 when it is there, and stop with a
    message that says so when it is not (no R in this container)."""
    from abc_dls_b200 import rabc
    from abc_dls_b200.engine import AbcError
    x = np.arange(20.0).reshape(10, 2)
    for call in (lambda: rabc.abc(target=[1.0, 2.0], param=x, sumstat=x, tol=0.5, method="neuralnet"),
                 lambda: rabc.abc(target=[1.0, 2.0], param=x, sumstat=x, tol=0.5, method="loclinear", transf="log"),
                 lambda: rabc.postpr(target=[1.0, 2.0], index=list("ab" * 5), sumstat=x, tol=0.5, method="mnlogistic"),
                 lambda: rabc.cv4abc(param=x, sumstat=x, nval=2, tols=0.5, method="neuralnet"),
                 lambda: rabc.cv4postpr(index=list("ab" * 5), sumstat=x, nval=2, tol=0.5, method="neuralnet")):
        with pytest.raises(AbcError, match="R's abc package could not be loaded through rpy2"):
            call()
    # with an R package object in place the call is forwarded untouched
    class FakeR:
        def abc(self, **kw):
            return ("R", sorted(kw))
    rabc._r_abc = FakeR()
    try:
        tag, keys = rabc.abc(target=[1.0], param=x[:, :1], sumstat=x[:, :1], tol=0.5, method="ridge")
        assert tag == "R" and {"target", "param", "sumstat", "tol", "method"} <= set(keys)
    finally:
        rabc._r_abc = None


def test_repeated_call_memo_is_safe():
    """abc() / postpr() reuse the validated views, the plan and the key of the previous call when they are made on the
    very same table objects (engine._last_call: weak references).  The memo must never serve another table, a changed
    target must be honoured, and the engine must not keep the caller's tables alive."""
    import gc
    import weakref
    from abc_dls_b200.engine import AbcEngine
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cpu_stages import CpuStages
    eng = AbcEngine(stages=CpuStages())
    par, ss, tgt = synth.abc_tables(5000, 3, 2, "cpu")
    a = eng.abc(tgt, par, ss, 0.02, "loclinear")
    assert eng._last_call is not None                                   # the second call will take the memo
    b = eng.abc(tgt, par, ss, 0.02, "loclinear")
    o = O.abc(tgt.numpy(), par.numpy().T, ss.numpy().T, 0.02, "loclinear")
    for r in (a, b):
        np.testing.assert_array_equal(r.idx.numpy(), o.idx)
        assert r.ds == o.ds
    tgt2 = tgt * 1.01                                                   # another target, same tables: memo + new result
    c = eng.abc(tgt2, par, ss, 0.02, "loclinear")
    o2 = O.abc(tgt2.numpy(), par.numpy().T, ss.numpy().T, 0.02, "loclinear")
    np.testing.assert_array_equal(c.idx.numpy(), o2.idx)
    assert c.ds == o2.ds and c.ds != a.ds
    c2 = eng.abc(tgt2, par, ss, 0.05, "rejection")                      # other tol / method: not the memoised plan
    assert c2.nacc == 250 and c2.adj_values is None
    with pytest.raises(Exception, match="Number of summary statistics"):
        eng.abc(tgt[:2], par, ss, 0.05, "rejection")                    # validation that depends on the target stays
    # other tables of the same shape (possibly at a recycled address / id): the memo must not serve them
    par3, ss3, tgt3 = synth.abc_tables(5000, 3, 2, "cpu", row_start=5000)
    d3 = eng.abc(tgt3, par3, ss3, 0.05, "rejection")
    o3 = O.abc(tgt3.numpy(), par3.numpy().T, ss3.numpy().T, 0.05, "rejection")
    np.testing.assert_array_equal(d3.idx.numpy(), o3.idx)
    # the engine holds weak references only
    w = weakref.ref(ss3)
    del ss3, d3
    gc.collect()
    assert w() is None, "the engine keeps the caller's table alive"
    # postpr: same rules
    ids, pss, ptgt = synth.postpr_tables(4000, "cpu")
    p1 = eng.postpr(ptgt, ids, ["A", "B", "C"], pss, 0.05)
    p2 = eng.postpr(ptgt, ids, ["A", "B", "C"], pss, 0.05)
    po = O.postpr(ptgt.numpy(), np.array(["A", "B", "C"])[ids.numpy()], pss.numpy().T, 0.05)
    for r in (p1, p2):
        np.testing.assert_array_equal(r.counts.numpy(), po.counts)
        np.testing.assert_array_equal(r.idx.numpy(), po.idx)
        assert abs(float(r.pred.sum()) - 1.0) < 1e-15 and r.mad.shape == (3,)
