"""Import shim (see ase/__init__.py): names pydbspm/SPMGrid.py imports at module scope; plotting only, off the hot path."""
import numpy as np

covalent_radii = np.full(119, 0.76)
