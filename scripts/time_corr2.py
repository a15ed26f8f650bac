"""Developer aid: time es / sr correlation + stages for one library build (DBSPM_B200_LIB), for a list of flag sets
(CORR_FLAGS="0,8192"), plus the cuFFT (torch.fft) correlation of the same shape as a yardstick."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth, _lib
FLAGS = [int(v) for v in os.environ.get("CORR_FLAGS", "0").split(",")]
SKIP = {"xfwd": 1 << 8, "yfwd": 1 << 9, "z": 1 << 10, "yinv": 1 << 11, "xinv": 1 << 12}; ALL = sum(SKIP.values())
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
dev = torch.device("cuda", 0)
tag = os.path.basename(_lib.LIB_PATH).replace("libdbspm_b200", "").replace(".so", "") or "default"
for name in (sys.argv[1:] or ["large_slab"]):
    cfg = synth.CONFIGS[name]; nx, ny, nz = cfg.shape
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g); B = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
    dr = np.array(cfg.cell) / np.array(cfg.shape); out = torch.empty((nx, ny, 54), dtype=torch.float64, device=dev)
    ref = None
    if os.environ.get("CORR_CUFFT", "1") == "1":
        def cufft():
            return torch.fft.irfftn(torch.fft.rfftn(A) * torch.conj(torch.fft.rfftn(B)), s=A.shape)[:, :, 76:130] * float(np.prod(dr))
        try:
            ref = cufft()
            t_cufft = timeit(cufft, reps=3, warm=1)
            print(f"CORR {tag:>10s} {name}: cuFFT rfftn x2 + product + irfftn + slice {t_cufft:.3f} ms", flush=True)
        except Exception as e:
            print("cufft comparator failed:", repr(e)[:200]); ref = None
        torch.cuda.empty_cache()
    for fl in FLAGS:
        t = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, out=out, flags=fl))
        err = float((out - ref).abs().max() / ref.abs().max()) if ref is not None else float("nan")
        tsr = timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, 1.08, 1.08, out=out, flags=fl))
        line = f"CORR {tag:>10s} {name} flags={fl:5d}: es {t:.3f} ms  sr {tsr:.3f} ms  err {err:.1e} |"
        for k, bit in SKIP.items():
            line += f" {k} {timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, out=out, flags=(ALL & ~bit) | fl)):.3f}"
        line += f" xfwd_pow {timeit(lambda: ops.density_overlap_window(A, B, dr, 76, 130, 1.08, 1.08, out=out, flags=(ALL & ~SKIP['xfwd']) | fl)):.3f}"
        print(line, flush=True)
    del ref
