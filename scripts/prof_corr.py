"""Developer aid: run one es correlation (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
name = sys.argv[1] if len(sys.argv) > 1 else "large_slab"
cfg = synth.CONFIGS[name]; dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
B = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
dr = np.array(cfg.cell) / np.array(cfg.shape)
for _ in range(2):
    out = ops.density_overlap_window(A, B, dr, 76, 130)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
