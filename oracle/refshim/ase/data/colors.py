"""Import shim (see ase/__init__.py)."""
import numpy as np

jmol_colors = np.full((119, 3), 0.5)
