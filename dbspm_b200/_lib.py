"""ctypes binding of libdbspm_b200.so (C ABI: include/dbspm_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load,
every entry point raises.  The oracle under oracle/ is test infrastructure and is never
imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DBSPM_B200_LIB selects another build of the same CUDA library (developer A/B timing of kernel variants)
LIB_PATH = os.environ.get("DBSPM_B200_LIB") or os.path.join(_HERE, "libdbspm_b200.so")

OK = 0
EINVAL, EUNSUPPORTED, EWORKSPACE, ECUDA = -1, -2, -3, -4

_lib = None


class DbspmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libdbspm_b200 error {code}: {message}")
        self.code = code


_vp = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "dbspm_version": (C.c_char_p, []),
    "dbspm_last_error": (C.c_char_p, []),
    "dbspm_corr3d_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "dbspm_corr3d_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "dbspm_corr3d_pruned_workspace_bytes": (C.c_size_t, [C.c_int] * 8),
    "dbspm_corr3d_pruned_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "dbspm_support_points_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "dbspm_corr3d_direct_workspace_bytes": (C.c_size_t, [C.c_int] * 6 + [C.POINTER(C.c_int)] * 3),
    "dbspm_corr3d_direct_cost": (C.c_longlong, [C.c_int] * 6 + [C.POINTER(C.c_int)] * 3),
    "dbspm_corr3d_direct_f64": (C.c_int, [_vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                          _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "dbspm_corr3d_dist_supported": (C.c_int, [C.c_int] * 4),
    "dbspm_corr3d_dist_fwd_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                            _vp, _vp, C.c_int, _vp]),
    "dbspm_corr3d_dist_fwd_peer_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.POINTER(_vp), C.c_int, _vp]),
    "dbspm_corr3d_pruned_geometry": (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_int)]),
    "dbspm_corr3d_dist_fwd_cut_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.POINTER(C.c_int), _vp, _vp, C.c_int, _vp]),
    "dbspm_corr3d_dist_fwd_peer_cut_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                                     C.c_int, C.POINTER(C.c_int), C.POINTER(_vp), C.c_int, _vp]),
    "dbspm_corr3d_dist_mid_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "dbspm_corr3d_dist_mid_peer_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, _vp, C.POINTER(_vp), C.c_int, _vp]),
    "dbspm_corr3d_dist_fin_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "dbspm_corr3d_dist_fin_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_size_t,
                                            C.c_int, _vp]),
    "dbspm_zsupport_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "dbspm_static_accumulate_f64": (C.c_int, [_vp, _vp, C.c_double, C.c_size_t, C.c_int, _vp]),
    "dbspm_relax_powell_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int,
                                         _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dbspm_relax_slab_halo": (C.c_int, [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "dbspm_relax_powell_slab_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double),
                                              C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int,
                                              _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dbspm_relax_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "dbspm_relax_workspace_bytes_layout": (C.c_size_t, [C.c_int] * 4),
    "dbspm_tricubic_gather_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, _vp, _vp]),
    "dbspm_dfz_gradient_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp]),
    "dbspm_convz_valid_f64": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_double, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C dbspm_b200/csrc`.  dbspm_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise DbspmError(rc, lib().dbspm_last_error().decode("utf-8", "replace"))


def version() -> str:
    return lib().dbspm_version().decode()
