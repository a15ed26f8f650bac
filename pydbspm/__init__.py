"""Import-compatible face of the reference package for the steps this repository implements
(sr / es / relax): `from pydbspm.interactions import get_density_overlap` etc. resolve to
dbspm_b200.  Modules of the reference that are outside that path (calculators, SPMGrid,
tools, sample) are not provided here."""
