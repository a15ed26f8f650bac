"""Developer aid: stage times of the distributed sr/es path for one rank of G (virtual ranks on one GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
dev = torch.device("cuda", 0)
name = sys.argv[1] if len(sys.argv) > 1 else "large_slab"
cfg = synth.CONFIGS[name]; nx, ny, nz = cfg.shape
g = torch.Generator(device=dev).manual_seed(1)
A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g); B = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
dr = np.array(cfg.cell) / np.array(cfg.shape); zl, zh = 76, 130
for G in (1, 2, 4, 8):
    if not ops.dist_supported(cfg.shape, G): print("G", G, "unsupported"); continue
    nxl = nx // G
    a, b = A[:nxl], B[:nxl]
    send = ops.dist_forward(a, b, G)
    t_f = timeit(lambda: ops.dist_forward(a, b, G, out=send))
    t_fp = timeit(lambda: ops.dist_forward(a, b, G, 1.08, 1.08, out=send))
    rA = torch.rand((nx, ny // G, nz), dtype=torch.float64, device=dev); rB = torch.rand_like(rA)
    W = ops.dist_middle(rA, rB, dr, cfg.shape, G, 0, zl, zh)
    t_m = timeit(lambda: ops.dist_middle(rA, rB, dr, cfg.shape, G, 0, zl, zh, out=W))
    Wall = torch.rand((G,) + tuple(W.shape), dtype=torch.float64, device=dev)
    out = ops.dist_finish(Wall, cfg.shape, G, zl, zh)
    t_e = timeit(lambda: ops.dist_finish(Wall, cfg.shape, G, zl, zh, out=out))
    print(f"{name} G={G}: fwd {t_f:.3f} (pow {t_fp:.3f})  mid {t_m:.3f}  fin {t_e:.3f} ms", flush=True)
    del send, rA, rB, W, Wall, out
