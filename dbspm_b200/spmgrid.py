"""`SPMGrid.interpolate_data` on the GPU (second half).

The reference's post-processing class gathers a data grid (e.g. the STM s/p-wave map) at the relaxed tip-apex positions of
another grid with one `tricubic.ip` call per point in a Python `nditer` loop (pydbspm/SPMGrid.py:311-342).  Here the index
arithmetic is restated verbatim on the host (same expressions, same order, same broadcasting -- and therefore the same
errors for shapes the reference cannot handle) and the interpolation is ONE launch of the tricubic gather kernel
(dbspm_tricubic_gather_f64).  The function is duck-typed: it accepts the reference's own SPMGrid objects or the small
container below, so a notebook can call it as `interpolate_data(tip_grid, stm_grid)` in place of
`tip_grid.interpolate_data(stm_grid)`.
"""
from __future__ import annotations

import copy

import numpy as np

from . import interactions as ops


class SPMGridLite:
    """The attributes of pydbspm.SPMGrid.SPMGrid that interpolate_data touches (SPMGrid.py:91-109): lateral coordinate
    meshes x, y (nx, ny), heights z, spacings dr, the data array, and -- for the grid that is interpolated AT -- the tip-apex
    displacements `tip` (nx, ny, nz, 3) with rpivot already added to the z component (set_tip, SPMGrid.py:292-309)."""

    def __init__(self, data, x, y, z, dr, tip=None, label="data"):
        self.label = label
        self.data = np.asarray(data, dtype=np.float64)
        self.x, self.y, self.z = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), np.asarray(z, dtype=np.float64)
        self.x_ = np.linalg.norm([self.x[:, 0], self.y[:, 0]], axis=0)
        self.y_ = np.linalg.norm([self.x[0, :], self.y[0, :]], axis=0)
        self.grid = [self.x_, self.y_, self.z]
        self.shape = self.data.shape
        self.dr = np.asarray(dr, dtype=np.float64)
        if tip is not None:
            self.tip = np.asarray(tip, dtype=np.float64)

    def set_tip(self, tip, rpivot=3.02, axis=0):
        """SPMGrid.set_tip: (3,nx,ny,nz) or (nx,ny,nz,3) displacements; rpivot is added to z (the apex, not the pivot)."""
        tip = np.array(tip, dtype=np.float64)
        if axis == 0:
            tip = np.moveaxis(tip, 0, -1)
        tip[..., 2] += rpivot
        self.tip = tip

    def copy(self):
        return copy.deepcopy(self)


def interpolate_data(self, data, device=None):
    """Drop-in for `SPMGrid.interpolate_data(self, data)` (SPMGrid.py:311-342): `data.data` interpolated at
    (grid point + tip displacement) of `self`; returns a copy of `self` whose `.data` holds the result."""
    field = ops.TricubicField(np.ascontiguousarray(data.data, dtype=np.float64), device=device)
    x_min, x_max = data.x_.min(), data.x_.max() + data.dr[0]
    y_min, y_max = data.y_.min(), data.y_.max() + data.dr[1]
    z_min, z_max = data.z.min(), data.z.max()
    idx = [(self.x_ >= x_min) & (self.x_ <= x_max),
           (self.y_ >= y_min) & (self.y_ <= y_max),
           (self.z >= z_min) & (self.z <= z_max)]
    tip_pos = self.tip
    tip_pos = tip_pos[idx[0], ...]
    tip_pos = tip_pos[:, idx[1], ...]
    tip_pos = tip_pos[..., idx[2], :]
    X = np.repeat(self.x[..., np.newaxis], len(self.z), axis=2) - x_min
    Y = np.repeat(self.y[..., np.newaxis], len(self.z), axis=2) - y_min
    Z = np.ones(self.shape) * self.z - z_min
    nX = (X + tip_pos[..., 0]) / data.dr[0]
    nY = (Y + tip_pos[..., 1]) / data.dr[1]
    nZ = (Z + tip_pos[..., 2]) / data.dr[2]
    res = field.gather(np.stack([nX, nY, nZ], axis=-1))
    new = self.copy()
    if new.shape != res.shape:
        new.shape = res.shape
        new.x = self.x[idx[0]]
        new.y = self.y[idx[1]]
        new.z = self.z[idx[2]]
        new.grid = [new.x, new.y, new.z]
    new.data = res
    return new
