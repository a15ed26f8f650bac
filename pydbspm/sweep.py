"""Alias: `pydbspm.sweep` resolves to dbspm_b200.sweep (B200 extension; no counterpart module in the reference)."""
from dbspm_b200.sweep import *  # noqa: F401,F403
from dbspm_b200.sweep import fdbm_sweep  # noqa: F401
