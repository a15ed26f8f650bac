"""Developer aid: one direct real-space correlation (pyridine grid, ball of radius 3 = 123 points) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS[os.environ.get("DIRECT_WORKLOAD", "large_slab")]; nx, ny, nz = cfg.shape
A = torch.rand(cfg.shape, dtype=torch.float64, device=dev)
ax = np.arange(-3, 4); gx, gy, gz = np.meshgrid(ax, ax, ax, indexing="ij"); m = gx ** 2 + gy ** 2 + gz ** 2 <= 9
pts = np.stack([gx[m] % nx, gy[m] % ny, gz[m] % nz], axis=1).astype(np.int32)
w = np.random.default_rng(0).random(len(pts)) + 0.1
dr = np.array(cfg.cell) / np.array(cfg.shape)
for _ in range(2):
    out = ops.density_overlap_direct(A, pts, w, dr, 76, 130, 1.08)
torch.cuda.synchronize(); print("ok", float(out.sum()))
