"""Drop-in alias: `pydbspm.setup_calculation` resolves to the B200 implementation (dbspm_b200.setup_calculation)."""
from dbspm_b200.setup_calculation import *  # noqa: F401,F403
from dbspm_b200 import setup_calculation as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
