"""Drop-in import face: the GPU form of SPMGrid.interpolate_data (pydbspm/SPMGrid.py:311-342 of the reference)."""
from dbspm_b200.spmgrid import SPMGridLite, interpolate_data  # noqa: F401
