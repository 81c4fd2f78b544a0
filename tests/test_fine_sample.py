"""The second-stage counting sample of the one-pass path (csrc/fused.cuh: k_sample_fine / k_fine_brackets) on
distributions that stress it: skew, heavy tails, ties, an atom at the median, rows sorted by a statistic (the sample is
stratified over the rows).  The brackets only decide WHAT the pass keeps; results stay R's bit for bit, and on
well-behaved tables the path must not miss."""
import numpy as np
import pytest
import torch

from oracle import abc_oracle as O

pytestmark = pytest.mark.gpu
N = (1 << 23) + 4242          # engine.FINE_MIN_ROWS = 2^23: the second stage is on


def _tables(kind, dev):
    g = torch.Generator(device="cpu").manual_seed(hash(kind) % 1000 + 7)
    z = torch.randn((3, N), generator=g, dtype=torch.float64)
    if kind == "lognormal":
        ss = torch.exp(0.8 * z)
    elif kind == "heavy_tails":
        ss = z / torch.rand((3, N), generator=g, dtype=torch.float64).clamp_min(1e-3)
    elif kind == "grid":
        ss = torch.round(z, decimals=3)                       # ~8e3 distinct values per column: ties everywhere
    elif kind == "sorted":
        ss = z.clone()
        ss[0] = torch.sort(z[0]).values                       # a trend along the rows
        ss[1] = z[1] + torch.linspace(-2, 2, N, dtype=torch.float64)
    elif kind == "atom":
        ss = z.clone()
        ss[2] = torch.where(torch.rand(N, generator=g) < 0.3, torch.zeros((), dtype=torch.float64), z[2])
    else:
        raise ValueError(kind)
    par = torch.stack([ss[0] * 0.5 + 0.1 * torch.randn(N, generator=g, dtype=torch.float64),
                       ss[1] - ss[2] + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)])
    tgt = torch.tensor([float(ss[0].median()), float(ss[1].median()) + 0.1, float(ss[2].median()) - 0.1],
                       dtype=torch.float64)
    return par.to(dev), ss.to(dev), tgt.to(dev)


@pytest.mark.parametrize("kind,must_hit", [("lognormal", True), ("heavy_tails", True), ("grid", False),
                                           ("sorted", False), ("atom", False)])
def test_fine_sample_keeps_r_exact(engine, kind, must_hit):
    par, ss, tgt = _tables(kind, engine.device)
    assert N >= engine.FINE_MIN_ROWS and engine.FINE_SHIFT > 0
    before, misses = dict(engine.path_calls), engine.fast_misses
    r = engine.abc(tgt, par, ss, 0.002, "loclinear")
    if must_hit:
        assert engine.path_calls["fused"] == before["fused"] + 1 and engine.fast_misses == misses
    o = O.abc(tgt.cpu().numpy(), par.cpu().numpy().T, ss.cpu().numpy().T, 0.002, "loclinear")
    np.testing.assert_array_equal(r.mad.cpu().numpy(), o.mad)
    assert r.ds == o.ds and r.nacc == o.nacc
    np.testing.assert_array_equal(r.idx.cpu().numpy(), o.idx)
    np.testing.assert_array_equal(r.unadj_values.cpu().numpy(), o.unadj_values)
    np.testing.assert_array_equal(r.weights.cpu().numpy(), o.weights)
    scale = np.abs(o.adj_values).max(axis=0, keepdims=True)
    assert float(np.max(np.abs(r.adj_values.cpu().numpy() - o.adj_values) / scale)) < 1e-9
    print(f"[fine:{kind}] path_calls={ {k: engine.path_calls[k] - before[k] for k in before} } "
          f"misses={engine.fast_misses - misses}")
