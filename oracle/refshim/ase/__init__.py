"""Import shim so the reference's pydbspm package (which imports ase at module
scope, pydbspm/grid.py:9-11) can be imported in a container without ase.
Only the names touched at import time are provided; none is on the hot path.
TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and oracle/ref_harness.py)."""
from .atoms import Atoms  # noqa: E402,F401  (pydbspm/SPMGrid.py: `from ase import Atoms`)
