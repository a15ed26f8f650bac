"""Drop-in alias: `pydbspm.parallel` resolves to the B200 implementation (dbspm_b200.parallel)."""
from dbspm_b200.parallel import *  # noqa: F401,F403
from dbspm_b200 import parallel as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
