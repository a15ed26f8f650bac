"""Drop-in alias: `pydbspm.calculate` resolves to the B200 implementation (dbspm_b200.calculate)."""
from dbspm_b200.calculate import *  # noqa: F401,F403
from dbspm_b200 import calculate as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
