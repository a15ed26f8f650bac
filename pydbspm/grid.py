"""Drop-in alias: `pydbspm.grid` resolves to the B200 implementation (dbspm_b200.grid)."""
from dbspm_b200.grid import *  # noqa: F401,F403
from dbspm_b200 import grid as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
