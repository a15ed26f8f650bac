"""Developer aid: the powered first pass (x_fwd with v**alpha fused) with pow_fast against libdevice exp(alpha*log v)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
SKIP = {"xfwd": 1 << 8, "yfwd": 1 << 9, "z": 1 << 10, "yinv": 1 << 11, "xinv": 1 << 12}; ALL = sum(SKIP.values())
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
dev = torch.device("cuda", 0)
for name in (sys.argv[1:] or ["pyridine", "large_slab"]):
    cfg = synth.CONFIGS[name]; nx, ny, nz = cfg.shape
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g) * 3.0
    B = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g) * 1e-3
    dr = np.array(cfg.cell) / np.array(cfg.shape); out = torch.empty((nx, ny, 54), dtype=torch.float64, device=dev)
    res = {}
    for label, fl in (("fast", 0), ("fast-wide", 1 << 7), ("slow", 1 << 6)):
        t_x = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, 1.08, 1.08, out=out, flags=(ALL & ~SKIP["xfwd"]) | fl))
        t = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, 1.08, 1.08, out=out, flags=fl))
        res[label] = out.clone()
        print(f"{name} {label}: sr {t:.3f} ms, x_fwd_pow {t_x:.3f} ms", flush=True)
    print(f"{name}: fast vs slow rel diff {float((res['fast'] - res['slow']).abs().max() / res['slow'].abs().max()):.2e}", flush=True)
