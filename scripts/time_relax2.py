"""Developer aid: time the relax kernel of one library build (DBSPM_B200_LIB selects it) on the bench potential
(static = V*sr + es + vdw of the synthetic large slab, cached in /dev/shm) and on synth.make_static_like.
   RELAX_LAYOUT = -1 (auto) | 0 | 1 | 2     RELAX_WORKLOAD = large_slab | tma_cu111 | pyridine"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth, _lib
from dbspm_b200.grid import Grid, get_grid_from_params
from dbspm_b200.input_parser import default_params
from argparse import Namespace
dev = torch.device("cuda", 0)
name = os.environ.get("RELAX_WORKLOAD", "large_slab")
layout = int(os.environ.get("RELAX_LAYOUT", "-1"))
cfg = synth.CONFIGS[name]
nx, ny, nz = cfg.shape
gS = Grid(cfg.shape, cell=np.diag(cfg.cell), positions=np.zeros((1, 3)), numbers=np.array([1]))
p = Namespace(**{**default_params, "Z0": cfg.z0, "ZF": cfg.zf, "ZREF": cfg.zref, "RPIVOT": cfg.rpivot, "ZPAD": cfg.zpad})
win, nidx = get_grid_from_params(p, gS, nidx=True, rpivot=cfg.rpivot, zpad=cfg.zpad)
_, _, nbox = get_grid_from_params(p, gS, nidx=True, nbox=True)
z_lo, z_hi = int(nidx[2].start), int(nidx[2].stop)
nz_relax = int(nbox[1, 2] - nbox[0, 2])
dr = gS.dr
cache = f"/dev/shm/static_{name}.pt"
if os.path.exists(cache):
    static = torch.load(cache).to(dev)
else:
    rhoS, potS, atoms = synth.make_sample(cfg, dev)
    rhoT, nrhoT = synth.make_tip(cfg, dev, "noisy")
    vdw = synth.make_vdw_window(cfg, atoms, (nx, ny, z_hi - z_lo), float(win.origin[2]), dev)
    sr = ops.density_overlap_window(rhoS, rhoT, dr, z_lo, z_hi, cfg.alpha, cfg.alpha)
    es = ops.density_overlap_window(potS, nrhoT, dr, z_lo, z_hi)
    static = ops.build_static([(sr, cfg.V), (es, 1.0), (vdw, 1.0)])
    del rhoS, potS, rhoT, nrhoT, sr, es, vdw
    ops.release_workspace(); torch.cuda.empty_cache()
    torch.save(static.cpu(), cache)
tag = os.path.basename(_lib.LIB_PATH).replace("libdbspm_b200", "").replace(".so", "") or "default"
for label, st in (("bench", static), ("synth", synth.make_static_like((nx, ny, z_hi - z_lo), 3, dev))):
    kw = {} if layout < 0 else {"layout": layout}
    r = ops.relax_grid(st, cfg.kappa, cfg.rpivot, dr, nz_relax, want_nfev=True, **kw)
    torch.cuda.synchronize()
    chk = float(r["E"].double().sum()), float(r["nfev"].float().mean())
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.relax_grid(st, cfg.kappa, cfg.rpivot, dr, nz_relax, **kw); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = float(np.median(ts))
    print(f"RELAX {tag:>14s} layout={layout:2d} {name} {label}: {t:8.2f} ms {nx*ny*nz_relax/t/1e3:8.1f} Mpos/s  sumE={chk[0]:.12e} nfev={chk[1]:.3f}", flush=True)
