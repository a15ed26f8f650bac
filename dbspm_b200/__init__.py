"""dbspm_b200 -- B200 (sm_100a) implementation of the data-parallel hot path of SPMTH/DBSPM:
the FDBM `sr` / `es` interaction grids and the `relax` step, behind the reference's own
entry points (`dbspm sr/es/relax -i input.in`, pydbspm.interactions / calculate / parallel)
and `.npz` layout.  Python host -> ctypes C ABI (include/dbspm_b200.h) -> hand-written CUDA.
"""
from . import _lib  # noqa: F401  (no import-time load: lib() raises when first used if missing)

__version__ = "0.1.0"
__all__ = ["interactions", "calculate", "parallel", "grid", "input_parser", "setup_calculation", "cli"]
