"""ctypes binding of the C ABI in ``include/abcdls.h`` (``libabcdls.so``, built in-tree by
``csrc/build.sh`` / ``__graft_entry__.build()``).

There is no fallback: if the shared library is missing the import of the engine fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libabcdls.so")

# status bits / codes (mirror include/abcdls.h)
OK, E_INVALID, E_WORKSPACE, E_CUDA, E_NODEVICE = 0, 1, 2, 3, 4
S_NONFINITE, S_FASTPATH_MISS, S_SINGULAR, S_ZERO_VAR, S_ZERO_VAR_REGION, S_CAPACITY = 1, 2, 4, 8, 16, 32
S_PEER_TIMEOUT, S_NONFINITE_FIT = 64, 128
SELECT_PASSES = 6
SELECT_BINS = 2048

_vp, _i32, _i64, _sz, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double

# name -> (restype, argtypes); every symbol declared in include/abcdls.h
SIGNATURES = {
    "abcdls_create": (_i32, [C.POINTER(_vp), _i32]),
    "abcdls_destroy": (_i32, [_vp]),
    "abcdls_last_error": (C.c_char_p, [_vp]),
    "abcdls_version": (_i32, []),
    "abcdls_sm_count": (_i32, [_vp]),
    "abcdls_select_workspace_bytes": (_sz, [_i32]),
    "abcdls_select_begin": (_i32, [_vp, _vp, _sz, _i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "abcdls_select_hist": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp]),
    "abcdls_select_pick": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "abcdls_select_hist_buffer": (_i32, [_vp, _vp, _i32, C.POINTER(_vp), C.POINTER(_sz)]),
    "abcdls_select_end": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "abcdls_median_from_ranks": (_i32, [_vp, _vp, _i32, _f64, _vp, _vp]),
    "abcdls_column_mad_workspace_bytes": (_sz, [_i32]),
    "abcdls_column_mad": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _sz, _vp, _vp]),
    "abcdls_scale_dist": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "abcdls_accept_workspace_bytes": (_sz, [_i64]),
    "abcdls_count_le": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "abcdls_accept": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "abcdls_gather_rows": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _vp]),
    "abcdls_fused_rows": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _sz, _vp, _vp]),
    "abcdls_fused_exact": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _sz,
                                  _vp, _vp]),
    "abcdls_gather_cand": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp,
                                  _vp, _vp, _vp, _vp, _vp]),
    "abcdls_fused_supported": (_i32, [_vp, _i64, _i32, _i64]),
    "abcdls_fused_workspace_bytes": (_sz, [_i64, _i32]),
    "abcdls_fused_sample_cols": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _i32, _vp, _sz, C.POINTER(_vp), C.POINTER(_sz),
                                        _vp]),
    "abcdls_fused_sample_fine": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _i32, _vp, _sz, C.POINTER(_vp), C.POINTER(_sz),
                                        _vp]),
    "abcdls_fused_sample_dist": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _sz, C.POINTER(_vp),
                                        C.POINTER(_sz), _vp]),
    "abcdls_fused_emit_cap": (_i64, [_i64]),
    "abcdls_fused_emit_pitch": (_i32, [_i32]),
    "abcdls_fused_pass": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _sz, _vp, C.POINTER(_vp),
                                 C.POINTER(_sz), _vp]),
    "abcdls_fused_refine": (_i32, [_vp, _vp, _i64, _i64, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp, _vp, _vp,
                                   C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_fused_candidates": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64,
                                       _vp, _vp, _vp, _sz, _vp, _vp]),
    "abcdls_fused_ds": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp, C.POINTER(_vp),
                               C.POINTER(_sz), _vp]),
    "abcdls_fused_check": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp, _vp]),
    "abcdls_gather": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp, _vp]),
    "abcdls_mad_fast_workspace_bytes": (_sz, [_i64, _i32]),
    "abcdls_mad_fast_sample": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _vp, _sz, C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_mad_fast_pass": (_i32, [_vp, _vp, _i64, _i32, _i64, _i32, _vp, _sz, C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_mad_fast_finish": (_i32, [_vp, _i64, _i64, _i32, _vp, _sz, _vp, _vp, _vp, _vp]),
    "abcdls_dist_fast_workspace_bytes": (_sz, [_i64, _i64]),
    "abcdls_dist_fast_sample": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _i64, _i64, _i32, _i64, _vp, _sz,
                                       C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_dist_fast_pass": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _i64, _i32, _i64, _i64, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _sz, _vp, C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_mad_fast_refine": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp, _vp, _vp,
                                      C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_dist_fast_refine": (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp,
                                       C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_gather_param_rows": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _vp]),
    "abcdls_normal_workspace_bytes": (_sz, [_i32, _i32]),
    "abcdls_normal": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "abcdls_solve": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "abcdls_residual": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "abcdls_residual_workspace_bytes": (_sz, [_i32]),
    "abcdls_residual_stats": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "abcdls_normal_logres": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "abcdls_recentre_log": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "abcdls_adjust": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "abcdls_col_stats_workspace_bytes": (_sz, [_i32]),
    "abcdls_col_stats": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "abcdls_resid_moments": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "abcdls_resid_finish": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "abcdls_final_small_doubles": (_sz, [_i32, _i32, _i32]),
    "abcdls_final_pack": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "abcdls_final_small": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                  _vp, _vp]),
    "abcdls_tail_supported": (_i32, [_i32, _i32]),
    "abcdls_tail_workspace_bytes": (_sz, [_vp, _i32, _i32]),
    "abcdls_tail_gather_fit": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64,
                                      _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "abcdls_tail_resid_fit": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz, _vp, _vp]),
    "abcdls_tail_adjust_final": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp,
                                        _vp, _vp, _sz, _vp, _vp, _vp]),
    "abcdls_columns_supported": (_i32, [_vp, _i64, _i32, _i64]),
    "abcdls_columns_workspace_bytes": (_sz, [_vp, _i64, _i32, _i64]),
    "abcdls_columns_sample": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _i64, _i64, _i64, _i32, _vp, _sz, C.POINTER(_vp),
                                     C.POINTER(_sz), _vp]),
    "abcdls_columns_pass": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _i64, _i32, _vp, _sz, C.POINTER(_vp), C.POINTER(_sz),
                                   C.POINTER(_vp), _vp]),
    "abcdls_columns_refine": (_i32, [_vp, _i64, _i64, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp, _vp, _vp,
                                     C.POINTER(_vp), C.POINTER(_sz), C.POINTER(_vp), C.POINTER(_sz), _vp]),
    "abcdls_columns_exact": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _sz, C.POINTER(_vp), C.POINTER(_i64), _vp, _vp,
                                    _vp]),
    "abcdls_columns_count": (_i32, [_vp, _i64, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp), _vp, _vp]),
    "abcdls_columns_rowkeys": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _vp, _sz, C.POINTER(_vp), _vp]),
    "abcdls_columns_accept": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp,
                                     _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "abcdls_wselect_workspace_bytes": (_sz, [_i32]),
    "abcdls_wselect": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "abcdls_col_moments": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "abcdls_density_workspace_bytes": (_sz, [_i32]),
    "abcdls_density_mode": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "abcdls_small_supported": (_i32, [_i64, _i64]),
    "abcdls_small_columns": (_i32, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp,
                                    _vp]),
    "abcdls_postpr_counts": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "abcdls_postpr_final": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp]),
    "abcdls_smc_workspace_bytes": (_sz, [_i64, _i32]),
    "abcdls_smc_oldrange": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _sz, _vp]),
    "abcdls_smc_box_filter": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _sz, _vp,
                                     _vp]),
    "abcdls_peer_handle_bytes": (_sz, []),
    "abcdls_peer_create": (_i32, [C.POINTER(_vp), _i32, _i32, _sz, _vp]),
    "abcdls_peer_connect": (_i32, [_vp, _vp]),
    "abcdls_peer_destroy": (_i32, [_vp]),
    "abcdls_peer_trace_read": (_i32, [_vp, _vp, _vp]),
    "abcdls_copy_out": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "abcdls_abc_single_workspace_bytes": (_sz, [_vp, _i64, _i32, _i32, _f64]),
    "abcdls_abc_single_nacc": (_i64, [_i64, _f64]),
    "abcdls_abc_single": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _f64, _i32, _i32, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "abcdls_peer_last_error": (C.c_char_p, [_vp]),
    "abcdls_peer_slot_bytes": (_sz, [_vp]),
    "abcdls_peer_allreduce": (_i32, [_vp, _vp, _sz, _i32, _i32, _vp]),
    "abcdls_peer_allgather": (_i32, [_vp, _vp, _sz, _vp, _vp]),
    "abcdls_peer_allgather_ragged": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
}

_lib = None


class AbcdlsLibraryError(ImportError):
    pass


def load() -> C.CDLL:
    """dlopen libabcdls.so and bind every declared symbol.  Raises if the library or a symbol is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AbcdlsLibraryError(
            f"{LIB_PATH} not found: build it with `bash abc-dls_b200/csrc/build.sh` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise AbcdlsLibraryError(f"symbol {name} missing from {LIB_PATH}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class AbcdlsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"abcdls error {code}: {msg}")
        self.code = code


class Handle:
    """Owns an abcdls_handle_t; ``call(name, *args)`` prepends the handle and checks the return code."""

    def __init__(self, device: int):
        self.lib = load()
        h = _vp()
        rc = self.lib.abcdls_create(C.byref(h), int(device))
        if rc != OK:
            raise AbcdlsError(rc, {E_NODEVICE: "no sm_100 (B200) CUDA device available — the engine has no CPU path",
                                   E_CUDA: "CUDA runtime error in abcdls_create"}.get(rc, "abcdls_create failed"))
        self.h = h
        self.sm_count = self.lib.abcdls_sm_count(h)

    def call(self, name: str, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != OK:
            raise AbcdlsError(rc, self.lib.abcdls_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.abcdls_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
