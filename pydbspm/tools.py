"""Drop-in alias: `pydbspm.tools` resolves to the B200 implementation (dbspm_b200.tools)."""
from dbspm_b200.tools import *  # noqa: F401,F403
from dbspm_b200.tools import get_fs  # noqa: F401
