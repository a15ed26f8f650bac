"""Developer aid: direct real-space correlation against the spectral path for ball-shaped tips of growing support
(measures the crossover and the fused-multiply-add rate the path choice in interactions.py uses)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
dev = torch.device("cuda", 0)
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for name in (sys.argv[1:] or ["pyridine", "large_slab"]):
    cfg = synth.CONFIGS[name]; nx, ny, nz = cfg.shape
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
    dr = np.array(cfg.cell) / np.array(cfg.shape); out = torch.empty((nx, ny, 54), dtype=torch.float64, device=dev)
    Nw = nx * ny * 54
    for radius in (1.0, 2.0, 3.0, 4.0, 5.0, 6.2, 7.5):
        r = int(np.ceil(radius))
        ax = np.arange(-r, r + 1)
        gx, gy, gz = np.meshgrid(ax, ax, ax, indexing="ij")
        m = gx ** 2 + gy ** 2 + gz ** 2 <= radius ** 2
        pts = np.stack([gx[m] % nx, gy[m] % ny, gz[m] % nz], axis=1).astype(np.int32)
        w = np.random.default_rng(0).random(len(pts)) + 0.1
        B = torch.zeros(cfg.shape, dtype=torch.float64, device=dev)
        B[torch.as_tensor(pts[:, 0].astype(np.int64)), torch.as_tensor(pts[:, 1].astype(np.int64)), torch.as_tensor(pts[:, 2].astype(np.int64))] = torch.as_tensor(w, device=dev)
        cost = ops.direct_cost(cfg.shape, 76, 130, pts)
        td = timeit(lambda: ops.density_overlap_direct(A, pts, w, dr, 76, 130, 1.0, out=out))
        d = out.clone()
        sup = ops.tip_zsupport(B)
        tf = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, out=out))
        err = float((out - d).abs().max() / out.abs().max())
        tp = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, out=out, tip_support=sup))
        tds = timeit(lambda: ops.density_overlap_direct(A, pts, w ** 1.08, dr, 76, 130, 1.08, out=out))
        print(f"DIRECT {name} K={len(pts):5d} fma/out={cost:5d} direct {td:7.3f} ms (sr-like {tds:7.3f}) | spectral {tf:7.3f} ms pruned {tp:7.3f} ms | "
              f"rate {Nw*cost/td/1e9:7.1f} GFMA/s  rel diff {err:.1e}  model picks direct: {ops.direct_is_faster(cfg.shape, 76, 130, pts)}", flush=True)
        del B
