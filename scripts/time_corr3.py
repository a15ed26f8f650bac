"""Developer aid: A/B of the axis-pass variants on the large slab (es: alpha = 1; sr: alpha = 1.08), stage times via the
SKIP flags, and a bit-identity check of the pipelined pass against the one-shot pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth, _lib
SKIP = {"xfwd": 1 << 8, "yfwd": 1 << 9, "z": 1 << 10, "yinv": 1 << 11, "xinv": 1 << 12}; ALL = sum(SKIP.values())
NO_PIPE = 1 << 14
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
    dr = np.array(cfg.cell) / np.array(cfg.shape)
    outs = {}
    for alpha in (1.0, 1.08):
        for label, fl in (("default", 0), ("nopipe", NO_PIPE), ("narrowy", 1 << 15), ("pownarrow", 1 << 16)):
            out = torch.empty((nx, ny, 54), dtype=torch.float64, device=dev)
            f = lambda extra=0: ops.density_overlap_window(A, B, dr, 76, 130, alpha, alpha, out=out, flags=fl | extra)
            t = timeit(f)
            outs[(alpha, label)] = out.clone()
            line = f"CORR3 {name} alpha={alpha} {label:9s}: total {t:.3f} ms |"
            for k, bit in SKIP.items():
                line += f" {k} {timeit(lambda: f(ALL & ~bit)):.3f}"
            f()
            print(line, flush=True)
        same = bool(torch.equal(outs[(alpha, "default")], outs[(alpha, "nopipe")]))
        print(f"CORR3 {name} alpha={alpha} default == nopipe bitwise: {same}; narrowy == nopipe: {bool(torch.equal(outs[(alpha, 'narrowy')], outs[(alpha, 'nopipe')]))}", flush=True)
