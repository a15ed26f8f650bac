"""Generate tests/golden/* by running the REFERENCE's own functions (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python oracle/make_golden.py

The reference package is imported unchanged from /root/reference with import shims for the
packages missing from this image (oracle/refshim: ase.atoms, ase.geometry, ase.units,
dftd3.interface -- none of them is reached on the sr/es/relax path).  The third-party
`tricubic` package is not installed and not vendored; the reference's relax_tip_z is driven
with oracle.Tricubic (our Lekien-Marsden restatement), so the relax goldens pin the
reference's control flow + scipy's Powell, and the tricubic boundary stays "parity unpinned".
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")

import pydbspm.grid as RG                      # noqa: E402  (the reference)
import pydbspm.input_parser as RP              # noqa: E402
import pydbspm.interactions as RI              # noqa: E402
import scipy                                   # noqa: E402

from oracle import oracle as O                 # noqa: E402
from dbspm_b200 import synth                   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
assert RI.__file__.startswith("/root/reference"), RI.__file__


def overlap_cases():
    rng = np.random.default_rng(20260100)
    data = {}
    meta = []
    for idx, (shape, alpha, signed) in enumerate([
        ((14, 15, 16), 1.0, True), ((14, 15, 16), 1.08, False), ((12, 10, 28), 1.08, False),
        ((9, 7, 10), 1.0, True), ((28, 30, 32), 1.0, True), ((16, 16, 16), 1.12, False),
    ]):
        a = rng.random(shape)
        b = rng.random(shape)
        if signed:
            b -= b.mean()                                 # zero-sum like nrhoT (grid.py:442-447)
        dr = rng.uniform(0.05, 0.2, 3)
        ref = RI.get_density_overlap(a ** alpha, b ** alpha, dr) if alpha != 1.0 else RI.get_density_overlap(a, b, dr)
        data[f"a{idx}"], data[f"b{idx}"], data[f"dr{idx}"], data[f"E{idx}"] = a, b, dr, ref
        meta.append({"idx": idx, "shape": list(shape), "alpha": alpha, "signed": signed})
    # a physically shaped case: synthetic 'tiny' config, sr and es on the real window
    cfg = synth.CONFIGS["tiny"]
    rhoS, potS, _ = synth.make_sample(cfg)
    rhoT, nrhoT = synth.make_tip(cfg, variant="noisy")
    dr = np.array(cfg.cell) / np.array(cfg.shape)
    sr = RI.get_density_overlap(rhoS.numpy() ** cfg.alpha, rhoT.numpy() ** cfg.alpha, dr)
    es = RI.get_density_overlap(potS.numpy(), nrhoT.numpy(), dr)
    data["tiny_sr_win"], data["tiny_es_win"], data["tiny_dr"] = sr[:, :, 20:44], es[:, :, 20:44], dr
    np.savez_compressed(os.path.join(OUT, "overlap_reference.npz"), **data)
    return meta


def relax_cases():
    static = synth.make_static_like((24, 20, 54), 11).numpy()
    dr = np.array([0.0765306, 0.0765306, 0.075])
    kappa, rp = 0.2, 3.02
    npivot = rp / dr[2]
    zi = np.arange(24)
    interp = O.Tricubic(list(static), list(static.shape))
    cols = [(0, 0), (5, 7), (12, 10), (11, 9), (23, 19), (7, 13), (17, 3), (2, 18)]
    ortho = np.array([RI.relax_tip_z(i, j, interp, kappa, npivot, zi, dr, "Powell") for i, j in cols])
    dR = np.diag(dr)
    noortho = np.array([RI.relax_noortho_tip_z(i, j, interp, kappa, npivot, zi, dr, dR, "Powell") for i, j in cols[:3]])
    stiff = np.array([RI.relax_tip_z(i, j, interp, 0.5, 2.0 / dr[2], np.arange(10), dr, "Powell") for i, j in cols[:3]])
    np.savez_compressed(os.path.join(OUT, "relax_reference.npz"), static=static, dr=dr, cols=np.array(cols),
                        ortho=ortho, noortho=noortho, stiff=stiff)
    return {"kappa": kappa, "rpivot": rp, "nz": 24, "stiff": {"kappa": 0.5, "rpivot": 2.0, "nz": 10}}


def relax_noortho_cases():
    """The reference's relax_noortho_tip_z (interactions.py:88-105) with a genuinely non-orthogonal step matrix dR
    (off-diagonal entries of a few per cent of the spacing: a monoclinic-like cell), which the orthorhombic examples
    never exercise.  Own file so that the fixtures of round 1 stay byte-identical."""
    static = synth.make_static_like((24, 20, 54), 11).numpy()
    dr = np.array([0.0765306, 0.0765306, 0.075])
    kappa, rp = 0.2, 3.02
    npivot = rp / dr[2]
    zi = np.arange(24)
    interp = O.Tricubic(list(static), list(static.shape))
    cols = [(0, 0), (5, 7), (12, 10), (11, 9), (23, 19), (7, 13)]
    dR = np.array([[dr[0], 0.011, 0.0], [0.004, dr[1], 0.0], [0.002, -0.003, dr[2]]])
    skew = np.array([RI.relax_noortho_tip_z(i, j, interp, kappa, npivot, zi, dr, dR, "Powell") for i, j in cols])
    np.savez_compressed(os.path.join(OUT, "relax_noortho_reference.npz"), dR=dR, cols=np.array(cols), skew=skew)
    return {"kappa": kappa, "rpivot": rp, "nz": 24, "static": "relax_reference.npz:static"}


def geometry_cases():
    out = {}
    for name, path, shape, cell in [
        ("pyridine", "/root/reference/example/input.in", (196, 196, 240), (15.0, 15.0, 18.0)),
        ("triangulene", "/root/reference/brstm_example/input.in", (240, 320, 240), (18.0, 24.0, 18.0)),
    ]:
        p = RP.parse_params(path)
        g = RG.Grid(shape, cell=np.diag(cell), positions=np.zeros((1, 3)), numbers=np.array([1]))
        win, nidx = RG.get_grid_from_params(p, g, nidx=True, rpivot=p.RPIVOT, zpad=p.ZPAD)
        rel, ridx, nbox = RG.get_grid_from_params(p, g, nidx=True, nbox=True)
        out[name] = {
            "params": {k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in vars(p).items()
                       if k in ("ALPHA", "V", "Z0", "ZF", "ZREF", "ZPAD", "RPIVOT", "K", "MINMETHOD", "COMPONENTS", "LABEL")},
            "dr": g.dr.tolist(),
            "window_shape": [int(v) for v in win.shape], "window_z": [int(nidx[2].start), int(nidx[2].stop)],
            "window_origin": [float(v) for v in win.origin], "window_span": np.asarray(win.span).tolist(),
            "relax_shape": [int(v) for v in rel.shape], "relax_z": [int(ridx[2].start), int(ridx[2].stop)],
            "relax_z_first": float(rel.z[0]), "relax_z_last": float(rel.z[-1]),
        }
    return out


def next_row_cases():
    """§8f rows: tricubic gather at relaxed tip positions (grid.py:325-361) and force -> Delta-f
    (tools.py:5-32), both run through the reference's own functions."""
    import types
    import pydbspm.tools as RT
    fake = types.ModuleType("tricubic")           # the reference does `import tricubic` inside the function
    fake.tricubic = O.Tricubic
    sys.modules["tricubic"] = fake
    rng = np.random.default_rng(20260101)
    cell = np.diag([3.6, 3.0, 4.5])
    data_grid = RG.Grid((24, 20, 30), cell=cell, positions=np.zeros((1, 3)), numbers=np.array([6]))
    x, y, z = np.meshgrid(np.arange(24), np.arange(20), np.arange(30), indexing="ij")
    sp = np.exp(-((x - 11.5) ** 2 + (y - 9.0) ** 2) / 40.0 - z / 8.0) + 0.05 * np.cos(2 * np.pi * x / 24) * np.sin(2 * np.pi * y / 20)
    data_grid.set_density("sp", sp)
    dr = data_grid.dr
    new_grid = RG.Grid(np.array([24, 20, 6]), cell=cell, span=np.diag([24, 20, 6]) * dr, origin=np.array([0.0, 0.0, 4 * dr[2]]),
                       zref=0.0, positions=np.zeros((1, 3)), numbers=np.array([6]))
    tip = rng.normal(0.0, 0.05, (3, 24, 20, 6))
    tip[2] = -0.3 + 0.02 * rng.standard_normal((24, 20, 6))
    for lab, arr in zip(("tip_x", "tip_y", "tip_z"), tip):
        new_grid.set_density(lab, arr)
    out = RG.interpolate_data(new_grid, data_grid, "sp", rpivot=0.3)
    F = rng.standard_normal((6, 5, 40))
    fs = RT.get_fs(F, dz=0.075, n=10)
    fs6 = RT.get_fs(F, dz=0.1, k0=1500.0, f0=25000.0, n=6)
    # vdW-style upsample (grid.py:122-139): coarse stride grid -> fine window grid
    coarse = RG.Grid(np.array([12, 10, 9]), cell=cell, span=np.diag([3.6, 3.0, 2.7]), origin=np.array([0.0, 0.0, 0.9]),
                     zref=0.0, positions=np.zeros((1, 3)), numbers=np.array([6]))
    cx, cy, cz = np.meshgrid(np.arange(12), np.arange(10), np.arange(9), indexing="ij")
    vdw_c = -1.0 / (1.0 + 0.05 * ((cx - 5.5) ** 2 + (cy - 4.5) ** 2) + 0.3 * cz)
    coarse.set_density("vdw", vdw_c)
    fine = RG.Grid(np.array([24, 20, 12]), cell=cell, span=np.diag([3.6, 3.0, 1.8]), origin=np.array([0.0, 0.0, 1.2]),
                   zref=0.0, positions=np.zeros((1, 3)), numbers=np.array([6]))
    fine.interpolate_density(coarse, "vdw")
    np.savez_compressed(os.path.join(OUT, "nextrow_reference.npz"), sp=sp, cell=cell, tip=tip, interp=out.data,
                        F=F, fs=fs, fs6=fs6, vdw_coarse=vdw_c, vdw_fine=fine.vdw)
    return {"interp_shape": list(out.data.shape), "rpivot": 0.3, "new_origin_z_index": 4}


def spmgrid_cases():
    """SPMGrid.interpolate_data (SPMGrid.py:311-342) run through the reference's own method on attribute containers (the
    class itself needs matplotlib / ase to be constructed; the method only reads attributes).  Own file."""
    import copy
    import types
    sys.path.insert(0, os.path.join(HERE, "refshim"))
    import pydbspm.SPMGrid as RS
    fake = types.ModuleType("tricubic")
    fake.tricubic = O.Tricubic
    sys.modules["tricubic"] = fake
    rng = np.random.default_rng(20260103)

    class Box(types.SimpleNamespace):
        def copy(self):
            return copy.deepcopy(self)

    def make(shape, dr, z0, data, tip=None):
        xs, ys = np.arange(shape[0]) * dr[0], np.arange(shape[1]) * dr[1]
        x, y = np.meshgrid(xs, ys, indexing="ij")
        b = Box(data=data, x=x, y=y, z=z0 + np.arange(shape[2]) * dr[2], dr=np.asarray(dr), shape=data.shape)
        b.x_ = np.linalg.norm([x[:, 0], y[:, 0]], axis=0)
        b.y_ = np.linalg.norm([x[0, :], y[0, :]], axis=0)
        if tip is not None:
            b.tip = tip
        return b
    dr = (0.15, 0.15, 0.1)
    gx, gy, gz = np.meshgrid(np.arange(24), np.arange(20), np.arange(30), indexing="ij")
    stm = np.exp(-((gx - 11.5) ** 2 + (gy - 9.0) ** 2) / 40.0 - gz / 8.0) + 0.05 * np.cos(2 * np.pi * gx / 24) * np.sin(2 * np.pi * gy / 20)
    data_grid = make((24, 20, 30), dr, 3.0, stm)
    tip = rng.normal(0.0, 0.05, (24, 20, 6, 3))
    tip[..., 2] = -3.02 + 0.02 * rng.standard_normal((24, 20, 6)) + 3.02       # set_tip adds rpivot to z
    tip_grid = make((24, 20, 6), dr, 3.4, np.zeros((24, 20, 6)), tip)
    out = RS.SPMGrid.interpolate_data(tip_grid, data_grid)
    np.savez_compressed(os.path.join(OUT, "spmgrid_reference.npz"), stm=stm, tip=tip, dr=np.asarray(dr), interp=out.data)
    return {"data_shape": [24, 20, 30], "tip_shape": [24, 20, 6], "data_z0": 3.0, "tip_z0": 3.4}


if __name__ == "__main__":
    if sys.argv[1:] == ["spmgrid"]:
        with open(os.path.join(OUT, "manifest.json")) as fh:
            manifest = json.load(fh)
        manifest["spmgrid"] = spmgrid_cases()
        with open(os.path.join(OUT, "manifest.json"), "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)
        print("wrote spmgrid_reference.npz")
        raise SystemExit(0)
    if sys.argv[1:] == ["noortho"]:          # round 2: add the skew-cell relax fixture without regenerating the others
        with open(os.path.join(OUT, "manifest.json")) as fh:
            manifest = json.load(fh)
        manifest["relax_noortho"] = relax_noortho_cases()
        with open(os.path.join(OUT, "manifest.json"), "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)
        print("wrote relax_noortho_reference.npz")
        raise SystemExit(0)
    manifest = {
        "generator": "oracle/make_golden.py", "reference": "/root/reference (SPMTH/DBSPM, unmodified)",
        "numpy": np.__version__, "scipy": scipy.__version__,
        "overlap": overlap_cases(), "relax": relax_cases(), "relax_noortho": relax_noortho_cases(), "spmgrid": spmgrid_cases(), "geometry": geometry_cases(),
        "nextrow": next_row_cases(),
    }
    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(OUT)))
