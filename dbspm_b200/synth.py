"""Deterministic synthetic inputs of the shapes BASELINE.json names.

The reference's real inputs (example/sample/sample.npz, tip/tip.npz) are git-LFS
pointers / absent, so parity and throughput are measured on synthetic densities with
the same shapes, dtypes and qualitative structure:

  rhoS   sum of Slater blobs Z_a (zeta^3/8pi) exp(-zeta |r - R_a|), lateral minimum image
  potS   sum of q_a erf(rho/0.6)/rho (signed)
  rhoT   two Slater blobs (O at the cell origin, C 1.153 A above it) wrapping around the
         cell corners; `compact` = multiplied by a C^2 cutoff (exact zero beyond 3.5 A),
         `noisy` = plus a 1e-7*max uniform floor (mimics CHGCAR noise)
  nrhoT  rhoT minus Gaussian cores rescaled to rhoT.sum() (zero net charge, grid.py:442-447)
  vdw    -sum C/(rho^6 + r0^6) on the window

Everything is float64 torch (CPU or CUDA); seeds are fixed per config.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

# pyridine atoms from example/sample/CONTCAR:9-19 (N, 5 C, 5 H), direct coordinates
_PYRIDINE_DIRECT = [
    (7, 0.44864, 0.39552667, 0.16666667), (6, 0.54059333, 0.55479333, 0.16666667),
    (6, 0.58760667, 0.47365333, 0.16666667), (6, 0.44681333, 0.55494, 0.16666667),
    (6, 0.53918667, 0.39326, 0.16666667), (6, 0.40140667, 0.47280667, 0.16666667),
    (1, 0.57468, 0.61383333, 0.16666667), (1, 0.65574, 0.47318, 0.16666667),
    (1, 0.41234, 0.61370667, 0.16666667), (1, 0.57152667, 0.33333333, 0.16666667),
    (1, 0.33333333, 0.47085333, 0.16666667),
]


@dataclass
class Config:
    name: str
    shape: tuple          # (Nx, Ny, Nz) of the sample grid
    cell: tuple           # orthorhombic cell lengths (A)
    seed: int
    z0: float = 2.7
    zf: float = 4.5
    zref: float = 3.0
    rpivot: float = 3.02
    zpad: float = 1.0
    alpha: float = 1.08
    V: float = 42.9071
    kappa: float = 0.2
    atoms: list = field(default_factory=list)   # (Z, x, y, z) cartesian; empty -> seeded random


def _pyridine_atoms():
    return [(z, a * 15.0, b * 15.0, c * 18.0) for (z, a, b, c) in _PYRIDINE_DIRECT]


CONFIGS = {
    # BASELINE.json configs[0]: example/input.in
    "pyridine": Config("pyridine", (196, 196, 240), (15.0, 15.0, 18.0), 20260100, atoms=_pyridine_atoms()),
    # configs[1]: brstm_example/input.in (Z0=2.5, ZF=5.0); atoms seeded (POSCAR not parsed here)
    "triangulene": Config("triangulene", (240, 320, 240), (18.0, 24.0, 18.0), 20260101, z0=2.5, zf=5.0),
    # configs[2]
    "tma_cu111": Config("tma_cu111", (384, 384, 160), (28.8, 28.8, 12.0), 20260102),
    # configs[4]
    "large_slab": Config("large_slab", (1024, 1024, 256), (76.8, 76.8, 19.2), 20260104),
    # small shapes for tests (same physics, scaled-down cell so dr stays ~0.075-0.15 A)
    "tiny": Config("tiny", (28, 30, 64), (4.2, 4.5, 9.6), 20260190, z0=0.6, zf=1.5, zref=3.0, rpivot=0.9, zpad=0.3),
}


def _seeded_atoms(cfg: Config, rng: np.random.Generator):
    lx, ly, _ = cfg.cell
    n_target = max(2, int(round(0.06 * lx * ly)))
    pts = []
    tries = 0
    while len(pts) < n_target and tries < 200 * n_target:
        tries += 1
        p = rng.uniform([0, 0], [lx, ly])
        ok = True
        for q in pts:
            d = np.abs(p - q[:2])
            d = np.minimum(d, [lx - d[0], ly - d[1]])
            if d[0] ** 2 + d[1] ** 2 < 1.3 ** 2:
                ok = False
                break
        if ok:
            pts.append(np.array([p[0], p[1]]))
    zs = rng.choice([1, 6, 6, 6, 7, 8], size=len(pts))
    return [(int(z), float(p[0]), float(p[1]), cfg.zref) for z, p in zip(zs, pts)]


def _axes(cfg: Config, device):
    nx, ny, nz = cfg.shape
    lx, ly, lz = cfg.cell
    x = torch.arange(nx, dtype=torch.float64, device=device) * (lx / nx)
    y = torch.arange(ny, dtype=torch.float64, device=device) * (ly / ny)
    z = torch.arange(nz, dtype=torch.float64, device=device) * (lz / nz)
    return x, y, z


def _min_image(d, L):
    d = torch.abs(d)
    return torch.minimum(d, L - d)


def make_sample(cfg: Config, device="cpu"):
    """Returns rhoS (>0), potS (signed), atoms list."""
    rng = np.random.default_rng(cfg.seed)
    atoms = cfg.atoms if cfg.atoms else _seeded_atoms(cfg, rng)
    charges = rng.uniform(-0.3, 0.3, size=len(atoms))
    x, y, z = _axes(cfg, device)
    lx, ly, _ = cfg.cell
    zeta = 2.2
    rho = torch.zeros(cfg.shape, dtype=torch.float64, device=device)
    pot = torch.zeros(cfg.shape, dtype=torch.float64, device=device)
    for (Z, ax, ay, az), q in zip(atoms, charges):
        dx2 = _min_image(x - ax, lx) ** 2
        dy2 = _min_image(y - ay, ly) ** 2
        dz2 = (z - az) ** 2
        r = torch.sqrt(dx2[:, None, None] + dy2[None, :, None] + dz2[None, None, :])
        rho += (Z * zeta ** 3 / (8 * math.pi)) * torch.exp(-zeta * r)
        rs = torch.clamp(r, min=1e-6)
        pot += float(q) * torch.erf(rs / 0.6) / rs
    return rho, pot, atoms


def make_tip(cfg: Config, device="cpu", variant="noisy"):
    """Returns rhoT (>=0), nrhoT (zero-sum).  Apex O atom at the cell origin (vasp.py:165-167)."""
    x, y, z = _axes(cfg, device)
    lx, ly, lz = cfg.cell
    zeta = 2.2
    rho = torch.zeros(cfg.shape, dtype=torch.float64, device=device)
    cores = torch.zeros(cfg.shape, dtype=torch.float64, device=device)
    rmin = None
    for Z, az in ((6.0, 0.0), (4.0, 1.153425)):
        dx2 = _min_image(x, lx) ** 2
        dy2 = _min_image(y, ly) ** 2
        dz2 = _min_image(z - az, lz) ** 2
        r = torch.sqrt(dx2[:, None, None] + dy2[None, :, None] + dz2[None, None, :])
        rho += (Z * zeta ** 3 / (8 * math.pi)) * torch.exp(-zeta * r)
        cores += Z * torch.exp(-(r / 0.35) ** 2)
        rmin = r if rmin is None else torch.minimum(rmin, r)
    if variant == "compact":
        rc = min(3.5, 0.45 * min(lx, ly, lz))
        t = torch.clamp(rmin / rc, max=1.0)
        rho = rho * (1 - t * t) ** 2          # C^1 at the cutoff, exact zero beyond it
        cores = cores * (1 - t * t) ** 2      # so that nrhoT is compact too (es takes the pruned path as well)
    elif variant == "noisy":
        gdev = device if torch.device(device).type == "cuda" else "cpu"   # big grids: draw on the GPU
        g = torch.Generator(device=gdev).manual_seed(cfg.seed + 7)
        noise = torch.rand(cfg.shape, dtype=torch.float64, generator=g, device=gdev)
        rho = rho + 1e-7 * float(rho.max()) * noise
    else:
        raise ValueError(variant)
    nrho = rho - cores * (rho.sum() / cores.sum())
    return rho, nrho


def make_vdw_window(cfg: Config, atoms, win_shape, z_origin, device="cpu"):
    """-sum_a C/(rho^6 + r0^6) evaluated on the window planes (z_origin = height of plane 0)."""
    x, y, _ = _axes(cfg, device)
    lx, ly, lz = cfg.cell
    dz = lz / cfg.shape[2]
    z = z_origin + torch.arange(win_shape[2], dtype=torch.float64, device=device) * dz
    out = torch.zeros(win_shape, dtype=torch.float64, device=device)
    for (_, ax, ay, az) in atoms:
        dx2 = _min_image(x - ax, lx) ** 2
        dy2 = _min_image(y - ay, ly) ** 2
        dz2 = (z - az) ** 2
        r2 = dx2[:, None, None] + dy2[None, :, None] + dz2[None, None, :]
        out -= 20.0 / (r2 ** 3 + 3.0 ** 6)
    return out


def make_static_like(shape, seed, device="cpu"):
    """A smooth, pyridine-like static potential on a window (for relax-only tests/benches):
    repulsive bumps + attractive wells decaying with height + a smooth corrugation."""
    nx, ny, nz = shape
    rng = np.random.default_rng(seed)
    x = torch.arange(nx, dtype=torch.float64, device=device)
    y = torch.arange(ny, dtype=torch.float64, device=device)
    z = torch.arange(nz, dtype=torch.float64, device=device)
    out = torch.zeros(shape, dtype=torch.float64, device=device)
    nb = max(3, (nx * ny) // 400)
    for _ in range(nb):
        cx, cy = rng.uniform(0, nx), rng.uniform(0, ny)
        amp = rng.uniform(0.1, 0.5) * (1 if rng.random() < 0.7 else -0.3)
        w = rng.uniform(12.0, 40.0)
        lam = rng.uniform(7.0, 20.0)
        dx = torch.abs(x - cx); dx = torch.minimum(dx, nx - dx)
        dy = torch.abs(y - cy); dy = torch.minimum(dy, ny - dy)
        out += amp * torch.exp(-(dx[:, None, None] ** 2 + dy[None, :, None] ** 2) / w) * torch.exp(-z / lam)[None, None, :]
    # smooth small-amplitude corrugation (band-limited: real DFT potentials have no grid-scale noise)
    for _ in range(4):
        kx, ky = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        ph = rng.uniform(0, 2 * math.pi)
        amp = rng.uniform(1e-3, 4e-3)
        out += amp * torch.cos(2 * math.pi * (kx * x[:, None, None] / nx + ky * y[None, :, None] / ny) + ph) \
            * torch.exp(-z / 15.0)[None, None, :]
    return out
