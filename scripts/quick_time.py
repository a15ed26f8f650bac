"""Developer timing aid (not the bench): per-stage CUDA-event timings of corr3d and relax."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth

SKIP = {"xfwd": 1 << 8, "yfwd": 1 << 9, "z": 1 << 10, "yinv": 1 << 11, "xinv": 1 << 12}
ALL = sum(SKIP.values())

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def main():
    dev = torch.device("cuda", 0)
    names = sys.argv[1:] or ["pyridine", "tma_cu111", "large_slab"]
    for name in names:
        cfg = synth.CONFIGS[name]
        nx, ny, nz = cfg.shape
        g = torch.Generator(device=dev).manual_seed(1)
        A = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
        B = torch.rand(cfg.shape, dtype=torch.float64, device=dev, generator=g)
        dr = np.array(cfg.cell) / np.array(cfg.shape)
        zlo, zhi = 76, 130
        out = torch.empty((nx, ny, zhi - zlo), dtype=torch.float64, device=dev)
        N = nx * ny * nz
        balg = 8 * (2 * N + nx * ny * (zhi - zlo))
        t_es = timeit(lambda: ops.density_overlap_window(A, B, dr, zlo, zhi, out=out))
        t_sr = timeit(lambda: ops.density_overlap_window(A, B, dr, zlo, zhi, 1.08, 1.08, out=out))
        t_v1 = timeit(lambda: ops.density_overlap_window(A, B, dr, zlo, zhi, out=out, flags=2))
        line = f"{name}: es {t_es:.3f} ms ({balg/t_es/1e6:.0f} GB/s alg) [axis v1: {t_v1:.3f}] sr {t_sr:.3f} ms ({balg/t_sr/1e6:.0f} GB/s alg) |"
        for k, bit in SKIP.items():
            t = timeit(lambda: ops.density_overlap_window(A, B, dr, zlo, zhi, out=out, flags=ALL & ~bit))
            line += f" {k} {t:.3f}"
        t = timeit(lambda: ops.density_overlap_window(A, B, dr, zlo, zhi, 1.08, 1.08, out=out, flags=ALL & ~SKIP['xfwd']))
        line += f" xfwd+pow {t:.3f}"
        print(line, flush=True)
        static = synth.make_static_like((nx, ny, 54), 3, dev)
        drr = dr
        t0 = time.time()
        t_rel = timeit(lambda: ops.relax_grid(static, 0.2, 3.02, drr, 24), reps=3, warm=1)
        r = ops.relax_grid(static, 0.2, 3.02, drr, 24, want_nfev=True)
        npos = nx * ny * 24
        print(f"{name}: relax {t_rel:.2f} ms  {npos/t_rel/1e3:.1f} Mpos/s  mean nfev {float(r['nfev'].float().mean()):.1f} max {int(r['nfev'].max())}", flush=True)
        del A, B, out, static, r
        ops.release_workspace(); torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
