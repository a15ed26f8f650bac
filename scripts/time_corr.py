"""Developer aid: time es correlation + stages for one library build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth, _lib
TMA = int(os.environ.get("CORR_FLAGS", "0"))
SKIP = {"xfwd": 1 << 8, "yfwd": 1 << 9, "z": 1 << 10, "yinv": 1 << 11, "xinv": 1 << 12}; ALL = sum(SKIP.values())
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
dev = torch.device("cuda", 0)
for name in (sys.argv[1:] or ["large_slab"]):
    cfg = synth.CONFIGS[name]; nx, ny, nz = cfg.shape
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g); B = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
    dr = np.array(cfg.cell) / np.array(cfg.shape); out = torch.empty((nx, ny, 54), dtype=torch.float64, device=dev)
    ref = torch.fft.irfftn(torch.fft.rfftn(A) * torch.conj(torch.fft.rfftn(B)), s=A.shape)[:, :, 76:130] * float(np.prod(dr)) if nx * ny * nz < 5e7 else None
    t = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, out=out, flags=TMA))
    err = float((out - ref).abs().max() / ref.abs().max()) if ref is not None else float("nan")
    line = f"{os.path.basename(_lib.LIB_PATH)} {name}: es {t:.3f} ms err {err:.1e} |"
    for k, bit in SKIP.items():
        line += f" {k} {timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, out=out, flags=(ALL & ~bit) | TMA)):.3f}"
    print(line, flush=True)
