"""The C-ABI shared library loads and exports every symbol include/abcdls.h declares (no compute)."""
import ctypes
import os
import re

from abc_dls_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "abcdls.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abcdls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in abcdls.h but not exported"


def test_binding_covers_header_and_version():
    assert sorted(_capi.SIGNATURES) == _declared()
    lib = _capi.load()
    assert lib.abcdls_version() == 1
    assert lib.abcdls_select_workspace_bytes(16) > 16 * 2 * 2048 * 8


def test_create_fails_loudly_without_a_b200():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(_capi.AbcdlsError):
        _capi.Handle(0)
    from abc_dls_b200 import rabc
    with pytest.raises(RuntimeError, match="no CPU path"):
        rabc.abc(target=[1.0], param=[[1.0], [2.0]], sumstat=[[1.0], [2.0]], tol=0.5, method="rejection")
