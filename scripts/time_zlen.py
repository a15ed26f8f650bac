"""Developer aid: per-stage time of the correlation as a function of the z length (1024 x 1024 x nz)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops
dev = torch.device("cuda", 0)
SKIP = {"x_fwd": 1 << 8, "y_fwd": 1 << 9, "z": 1 << 10, "y_inv": 1 << 11, "x_inv": 1 << 12}
nxy = int(os.environ.get("NXY", "1024"))
for nz in [int(v) for v in sys.argv[1:]]:
    A = torch.rand((nxy, nxy, nz), dtype=torch.float64, device=dev)
    B = torch.rand((nxy, nxy, nz), dtype=torch.float64, device=dev)
    out = torch.empty((nxy, nxy, 54), dtype=torch.float64, device=dev)
    st = {}
    for nm in ("x_fwd", "y_fwd", "z"):
        ts = []
        for it in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.density_overlap_window(A, B, np.ones(3), 1, 55, out=out, flags=sum(SKIP.values()) & ~SKIP[nm]); b.record(); torch.cuda.synchronize()
            if it: ts.append(a.elapsed_time(b))
        st[nm] = round(float(np.mean(ts)), 3)
    print(f"nz={nz} n={nz//2}: {st} total3={sum(st.values()):.3f}", flush=True)
    del A, B
