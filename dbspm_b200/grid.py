"""Scan-window geometry and the `.npz` layout (host side, numpy only).

Mirrors the semantics of the reference's pydbspm/grid.py so that files written here are
readable by the reference's `read_grid` / `SPMGrid` and vice versa:

  Grid                   pydbspm/grid.py:19-156   (attributes x, y, z, dr, dR, ortho, lvec, ...)
  lattice_axes           pydbspm/grid.py:179-202  (get_grid)
  get_grid_from_params   pydbspm/grid.py:205-273  (window from Z0/ZF/ZREF/RPIVOT/ZPAD)
  save_grid / read_grid  pydbspm/grid.py:159-176, :306-323
  sph2grid / sph2cart    pydbspm/grid.py:364-386

Only what the sr / es / relax steps touch is implemented (no ASE `atoms`, no `reduce`
stride grid -- those belong to the DFT / vdw steps that stay with the reference).
"""
from __future__ import annotations

import copy

import numpy as np

_META_KEYS = ("shape", "span", "origin", "ortho", "cell", "numbers", "positions", "zref")


def _angle(u, v) -> float:
    return float(np.arccos(np.dot(u, v) / (np.linalg.norm(u) * np.linalg.norm(v))))


def lattice_axes(cell, shape, mesh=None, origin=None):
    """Coordinates and spacings of a regular grid on `cell` (get_grid, grid.py:179-202).

    Returns x, y, z, dr, dR, ortho, lvec.  The z axis must be orthogonal to x and y."""
    cell = np.asarray(cell, dtype=float)
    shape = np.asarray(shape)
    right = np.pi / 2
    if _angle(cell[0], cell[2]) != right or _angle(cell[2], cell[1]) != right:
        print(f"Cell is not orthogonal: {cell}")
        raise ValueError("z-axis is not orthogonal to xy-axes.")
    ortho = 1 if _angle(cell[0], cell[1]) == right else 0
    lvec = np.linalg.norm(cell, axis=1)
    dr = lvec / shape
    dR = cell / shape
    if ortho and not mesh:
        x = np.arange(shape[0]) * dr[0]
        y = np.arange(shape[1]) * dr[1]
    else:
        frac = np.indices(tuple(shape[:2]), dtype=float).T / shape[:2]
        x, y = np.dot(frac, cell[:2, :2]).T
    z = np.arange(shape[2]) * dr[2]
    if origin is not None:
        x = x + origin[0]
        y = y + origin[1]
        z = z + origin[2]
    return x, y, z, dr, dR, ortho, lvec


get_grid = lattice_axes  # reference name


class Grid:
    """A regular grid (full cell or a z-window of it) plus named 3-D arrays ("densities")."""

    def __init__(self, shape, cell=None, span=None, origin=None, zref=None, z0=None, zf=None,
                 atoms=None, positions=None, numbers=None, verbose=False, **_ignored):
        if atoms is not None:
            raise NotImplementedError("Grid(atoms=...) needs ASE and belongs to the DFT steps of the reference")
        if cell is None and span is None:
            raise ValueError("One of cell or span must be defined.")
        self.cell = np.array(span if cell is None else cell)
        self.shape = shape
        self.zref = None
        if (origin is None) != (span is None):
            raise ValueError("Both origin and span must be defined.")
        if origin is not None:
            self.origin, self.span = origin, span
            if zref is not None:
                self.zref = zref
                self.z0 = self.origin[2] - self.zref
                self.zf = self.z0 + self.span[2]
        elif (zref is None) != (z0 is None) or (z0 is None) != (zf is None):
            raise ValueError("All of zref, z0, and zf values must be defined.")
        elif zref is not None:
            self.zref, self.z0, self.zf = zref, z0, zf
            self.origin = np.array([0.0, 0.0, zref + z0])
            c = np.asarray(cell, dtype=float)
            self.span = np.array([c[0], c[1], [0.0, 0.0, zf - z0]])
        else:
            self.origin = np.zeros(3)
            self.span = self.cell
            self.zref = 0.0
        self.x, self.y, self.z, self.dr, self.dR, self.ortho, self.lvec = lattice_axes(
            self.span, self.shape, mesh=True, origin=self.origin)
        self.z = self.z - self.zref
        if positions is not None and numbers is not None:
            assert positions.shape[0] == len(numbers)
            self.positions, self.numbers = positions, numbers
        self.labels = []

    # -- named arrays ---------------------------------------------------------------------
    def set_density(self, label: str, data, safe=True):
        if safe and label in self.__dict__:
            print(f"Density labeled {label} already exists. Choose a different label or set `safe=False`")
        assert data.ndim == 3, "Data array is not 3D."
        setattr(self, label, data)
        self.labels.append(label)

    def save(self, filename: str):
        save_grid(filename, self, **{k: getattr(self, k) for k in self.labels})

    def save_density(self, *labels: str, filename: str):
        save_grid(filename, self, **{k: getattr(self, k) for k in labels})

    def set_zref(self, zref: float):
        self.z = (self.z + self.zref) - zref
        self.zref = zref

    def interpolate_density(self, grid, label):
        """Tricubic resampling of `grid.<label>` onto this grid's points (pydbspm/grid.py:122-139,
        used at dbspm:458 to bring the stride-2 vdW grid back to the full window): one gather kernel
        launch over all points instead of a Python nditer loop over `interp.ip`."""
        from .interactions import TricubicField

        field = TricubicField(np.ascontiguousarray(getattr(grid, label), dtype=np.float64))
        nz = len(self.z)
        X = np.repeat(self.x[..., np.newaxis], nz, axis=2) - grid.x.min()
        Y = np.repeat(self.y[..., np.newaxis], nz, axis=2) - grid.y.min()
        Z = np.ones(tuple(int(v) for v in self.shape)) * self.z - grid.z.min()
        pts = np.stack([X / grid.dr[0], Y / grid.dr[1], Z / grid.dr[2]], axis=-1)
        self.set_density(label, field.gather(pts))

    def copy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return f"Grid({self.cell}, {self.shape}, {self.dr})"


def save_grid(filename, grid: Grid, **arrays):
    """Uncompressed .npz: the data arrays plus the eight geometry keys (grid.py:159-176)."""
    for v in arrays.values():
        assert np.equal(v.shape, grid.shape).all(), "Shape mismatch."
    meta = {k: getattr(grid, k, None) for k in _META_KEYS}
    with open(filename, "wb") as fh:
        np.savez(fh, **arrays, **meta)


def read_grid(filename):
    """Inverse of save_grid (grid.py:306-323, without the ASE branch)."""
    with np.load(filename, allow_pickle=True) as f:
        g = Grid(shape=f["shape"], span=f["span"], origin=f["origin"], cell=f["cell"],
                 numbers=f["numbers"], positions=f["positions"], zref=f["zref"])
        for key in f.files:
            if key not in _META_KEYS:
                g.set_density(key, f[key])
    return g


def get_grid_from_params(params, grid: Grid, nidx=False, nbox=False, rpivot=0.0, zpad=0.0,
                         numbers=None, positions=None, reduce=False):
    """Scan window derived from Z0/ZF/ZREF (+ RPIVOT/ZPAD head-room), grid.py:205-273.

    z-extent: [Z0, ZF + zpad] if ZF - Z0 >= rpivot else [Z0, Z0 + rpivot + zpad], shifted by
    ZREF; index box = rint(bbox / dr) with the top z index forced to bottom + rint(span/dz).
    Side effect kept from the reference: params.BBOX is filled in on first use."""
    if reduce:
        raise NotImplementedError("stride-reduced grids belong to the reference's vdw step")
    if params.BBOX is None:
        params.BBOX = np.array([[0, 0, params.Z0], [grid.lvec[0], grid.lvec[1], params.ZF]])
    box = params.BBOX.copy()
    if box[1, 2] - box[0, 2] >= rpivot:
        box[1, 2] += zpad
    else:
        box[1, 2] = box[0, 2] + rpivot + zpad
    box[:, 2] += params.ZREF
    span = np.diag(box[1] - box[0]).astype(float)
    nspan = np.rint(span / grid.dr).astype(int)
    ibox = np.rint(box / grid.dr).astype(int)
    ibox[1, 2] = ibox[0, 2] + nspan[2, 2]
    window = Grid(nspan.diagonal(), cell=grid.cell, span=nspan * grid.dr,
                  origin=np.rint(box[0] / grid.dr) * grid.dr, zref=params.ZREF,
                  numbers=numbers, positions=positions)
    slices = tuple(slice(lo, hi) for lo, hi in zip(ibox[0], ibox[1]))
    if nidx and nbox:
        return window, slices, ibox
    if nidx:
        return window, slices
    if nbox:
        return window, ibox
    return window


def sph2grid(th, az, r, dr):
    """(theta, phi, r) -> offset in grid-index units (grid.py:364-369)."""
    return [r * np.sin(th) * np.cos(az) / dr[0], r * np.sin(th) * np.sin(az) / dr[1], r * np.cos(th) / dr[2]]


def sph2cart(th, az, r):
    """(theta, phi, r) -> Cartesian (grid.py:381-386)."""
    return r * np.sin(th) * np.cos(az), r * np.sin(th) * np.sin(az), r * np.cos(th)


def interpolate_data(grid_new, grid_data, label, rpivot=3.02):
    """Tricubic gather of `grid_data.<label>` at the relaxed tip-apex positions of `grid_new`
    (pydbspm/grid.py:325-361, the BRSTM step dbspm:655-664): every point (x,y,z) of grid_new is
    displaced by (tip_x, tip_y, tip_z + rpivot), converted to index units of grid_data and
    interpolated -- one gather kernel launch instead of a Python nditer loop.

    Points of grid_new outside grid_data's extent are dropped on all operands (the reference
    only masks the tip offsets, which fails unless the mask is all-True)."""
    from .interactions import TricubicField

    field = TricubicField(np.ascontiguousarray(getattr(grid_data, label), dtype=np.float64))

    def lateral(g):
        return (np.linalg.norm([g.x[:, 0], g.y[:, 0]], axis=0), np.linalg.norm([g.x[0, :], g.y[0, :]], axis=0))

    dx_, dy_ = lateral(grid_data)
    grid_new.x_, grid_new.y_ = lateral(grid_new)
    x_min, x_max = dx_.min(), dx_.max() + grid_data.dr[0]
    y_min, y_max = dy_.min(), dy_.max() + grid_data.dr[1]
    z_min, z_max = grid_data.z.min(), grid_data.z.max()
    keep = [(grid_new.x_ >= x_min) & (grid_new.x_ <= x_max),
            (grid_new.y_ >= y_min) & (grid_new.y_ <= y_max),
            (grid_new.z >= z_min) & (grid_new.z <= z_max)]
    sel = np.ix_(*keep)
    nz = len(grid_new.z)
    X = np.repeat(grid_new.x[..., np.newaxis], nz, axis=2) - x_min
    Y = np.repeat(grid_new.y[..., np.newaxis], nz, axis=2) - y_min
    Z = np.ones(tuple(int(v) for v in grid_new.shape)) * grid_new.z - z_min
    pts = np.stack([(X[sel] + grid_new.tip_x[sel]) / grid_data.dr[0],
                    (Y[sel] + grid_new.tip_y[sel]) / grid_data.dr[1],
                    (Z[sel] + grid_new.tip_z[sel] + rpivot) / grid_data.dr[2]], axis=-1)
    res = field.gather(pts)
    new = copy.deepcopy(grid_new)
    if tuple(int(v) for v in new.shape) != res.shape:
        new.shape = np.array(res.shape)
        new.x, new.y, new.z = grid_new.x[keep[0]], grid_new.y[keep[1]], grid_new.z[keep[2]]
        new.grid = [new.x, new.y, new.z]
    new.set_density("data", res, safe=False)
    return new
