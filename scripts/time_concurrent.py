"""Developer aid: sr and es of the large slab one after the other on one stream against both at once on two streams."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "large_slab"]; nx, ny, nz = cfg.shape
g = torch.Generator(device=dev).manual_seed(1)
A, B, Cc, D = (torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g) for _ in range(4))
dr = np.array(cfg.cell) / np.array(cfg.shape)
o1 = torch.empty((nx, ny, 54), dtype=torch.float64, device=dev); o2 = torch.empty_like(o1)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def seq():
    ops.density_overlap_window(A, B, dr, 76, 130, 1.08, 1.08, out=o1)
    ops.density_overlap_window(Cc, D, dr, 76, 130, out=o2)
def conc():
    cur = torch.cuda.current_stream(dev)
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): ops.density_overlap_window(A, B, dr, 76, 130, 1.08, 1.08, out=o1)
    with torch.cuda.stream(s2): ops.density_overlap_window(Cc, D, dr, 76, 130, out=o2)
    cur.wait_stream(s1); cur.wait_stream(s2)
def timeit(fn, reps=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
seq(); r1, r2 = o1.clone(), o2.clone()
print(f"CONC sequential {timeit(seq):.3f} ms", flush=True)
print(f"CONC two streams {timeit(conc):.3f} ms", flush=True)
conc(); torch.cuda.synchronize()
print("CONC same bits:", bool(torch.equal(r1, o1) and torch.equal(r2, o2)), flush=True)
