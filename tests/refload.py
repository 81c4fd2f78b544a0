"""Loads the UNMODIFIED reference (``/root/reference/src``) in this container so that its own Python can serve as
the oracle / integration driver of the boundary.

The reference's ABC step needs R + rpy2, and its modules import tensorflow / h5py / allel / tzlocal at the top
(``Classes/ABC.py:13-37``, ``SFS/Class.py:12``) - none of which exist here.  Those imports are satisfied with inert
stand-ins in ``sys.modules`` (nothing on the tested paths touches them), then the two module globals the hot path goes
through are rebound exactly as ``INTEGRATION.md`` tells a maintainer to:

    ABC.abc      = abc_dls_b200.rabc        (was: Misc.importr_tryhard('abc'), ABC.py:36)
    ABC.robjects = abc_dls_b200.rshim       (was: rpy2.robjects, ABC.py:17)

``/root/reference`` does not exist on the GPU box: every user of this module skips when ``available()`` is False, and
what has to travel is committed as fixtures under ``tests/golden/`` by ``tests/golden/make_ref_golden.py``.
"""
import os
import sys
import types

REF_SRC = "/root/reference/src"
_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "Classes", "ABC.py"))


class _Inert:
    """Accepts any construction / call / attribute: stands in for Keras layers, scalers' type hints, etc."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """-> (ABC module, SFS.Class module) of the reference, bound to this repo's rabc / rshim."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("the reference tree is not present (GPU box): use the committed fixtures")
    for n in ("h5py", "tzlocal", "allel"):
        if n not in sys.modules:
            _stub(n)
    if "tensorflow" not in sys.modules:
        tf = _stub("tensorflow")
        tf.config, tf.keras = _Inert(), _Inert()
        _stub("tensorflow.keras")
        _stub("tensorflow.keras.models", Sequential=_Inert, Model=_Inert)
        _stub("tensorflow.keras.layers", Dense=_Inert, GaussianNoise=_Inert, Dropout=_Inert)
        _stub("tensorflow.keras.callbacks", EarlyStopping=_Inert, ModelCheckpoint=_Inert, ReduceLROnPlateau=_Inert)
    if "rpy2" not in sys.modules:
        rpy2 = _stub("rpy2")
        ro = _stub("rpy2.robjects")
        rpy2.robjects = ro
        ro.pandas2ri = _stub("rpy2.robjects.pandas2ri", activate=lambda: None)
        ro.packages = _stub("rpy2.robjects.packages", importr=lambda name: _Inert())
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from Classes import ABC          # noqa: E402  (the reference's module, unmodified)
    from SFS import Class as SFSClass  # noqa: E402
    from abc_dls_b200 import rabc, rshim
    ABC.abc = rabc
    ABC.robjects = rshim
    _loaded = (ABC, SFSClass)
    return _loaded
