"""Drop-in alias: `pydbspm.input_parser` resolves to the B200 implementation (dbspm_b200.input_parser)."""
from dbspm_b200.input_parser import *  # noqa: F401,F403
from dbspm_b200 import input_parser as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
