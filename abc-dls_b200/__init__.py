"""abc_dls_b200 — B200-native engine for the ABC acceptance-and-adjustment step of ABC-DLS.

Import as ``abc_dls_b200`` (the directory is named ``abc-dls_b200``; ``abc_dls_b200/__init__.py`` at the
repo root aliases it).  ``rabc`` is the drop-in for the rpy2 ``abc`` handle; ``engine.AbcEngine`` works on
device tables; everything numeric is CUDA (sm_100a) behind the C ABI in ``include/abcdls.h``.
"""
from . import _capi  # noqa: F401  (loading the shared library is deferred to first use)

__all__ = ["rabc", "engine", "cv", "smc", "synth", "summary", "rshim", "stages", "comm"]
__version__ = "0.1.0"
