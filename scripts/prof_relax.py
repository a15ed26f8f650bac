"""Developer aid: two relax launches on the bench potential (cached by scripts/time_relax2.py) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops
dev = torch.device("cuda", 0)
name = os.environ.get("RELAX_WORKLOAD", "large_slab")
layout = int(os.environ.get("RELAX_LAYOUT", "-1"))
static = torch.load(f"/dev/shm/static_{name}.pt").to(dev)
dr = np.array([0.075, 0.075, 0.075])
kw = {} if layout < 0 else {"layout": layout}
for _ in range(2):
    r = ops.relax_grid(static, 0.2, 3.02, dr, 24, **kw)
torch.cuda.synchronize()
print("ok", float(r["E"].sum()))
