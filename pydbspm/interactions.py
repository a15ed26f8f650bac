"""Drop-in alias: `pydbspm.interactions` resolves to the B200 implementation (dbspm_b200.interactions)."""
from dbspm_b200.interactions import *  # noqa: F401,F403
from dbspm_b200 import interactions as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
