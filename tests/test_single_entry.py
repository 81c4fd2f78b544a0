"""The call-level C entry point (csrc/single.cu: abc() on one GPU as ONE C call) against the engine, which sequences
the same stages from Python, and against the oracle."""
import numpy as np
import pytest
import torch

from oracle import abc_oracle as O
from abc_dls_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,P,tol,method,layout", [((1 << 20) + 64, 4, 3, 0.01, "loclinear", "row"),
                                                       ((1 << 19) + 2, 5, 2, 0.004, "rejection", "col"),
                                                       ((1 << 20) + 130, 19, 4, 0.01, "loclinear", "row")])
def test_call_level_entry_matches_engine_and_oracle(engine, n, d, P, tol, method, layout):
    par, ss, tgt = synth.abc_tables(n, d, P, engine.device)
    if layout == "row":
        par = par.t().contiguous().t()                              # (P, n) view of an n x P C-order table
    out = engine.k.abc_single(tgt, par, ss, tol, method)
    torch.cuda.synchronize()
    small = out["small"].cpu().numpy()
    assert int(small[0]) == 0, f"status bits {int(small[0])}"
    n_acc = int(small[1])
    o = O.abc(tgt.cpu().numpy(), par.cpu().numpy().T, ss.cpu().numpy().T, tol, method)
    assert n_acc == o.nacc == out["nacc"] and small[2] == o.ds
    np.testing.assert_array_equal(out["idx"][:n_acc].cpu().numpy(), o.idx)
    np.testing.assert_array_equal(small[8:8 + d], o.mad)
    np.testing.assert_array_equal(out["y"][:, :n_acc].t().cpu().numpy(), o.unadj_values)
    np.testing.assert_array_equal(out["ss_out"][:, :n_acc].t().cpu().numpy(), o.ss)
    r = engine.abc(tgt, par, ss, tol, method)
    assert torch.equal(r.idx, out["idx"][:n_acc])
    if method == "loclinear":
        np.testing.assert_array_equal(out["w"][:n_acc].cpu().numpy(), o.weights)
        adj = out["adj"][:, :n_acc].t().cpu().numpy()
        scale = np.abs(o.adj_values).max(axis=0, keepdims=True)
        assert float(np.max(np.abs(adj - o.adj_values) / scale)) < 1e-9
        assert float(np.max(np.abs(adj - r.adj_values.cpu().numpy()) / scale)) < 1e-12
