"""Developer aid: run the CPU emulation of the kernel bodies under AddressSanitizer (index arithmetic of the row-mapped,
peer-store, z-pruned and TMA-emulated passes).  Build and run:
    g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fsanitize=address -shared -o /tmp/libdbspm_emu_asan.so \
        tests/emu/emu.cpp dbspm_b200/csrc/plan.cpp -lm
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python scripts/asan_emulation.py"""
import ctypes as C, numpy as np, sys
L = C.CDLL(sys.argv[1] if len(sys.argv) > 1 else "/tmp/libdbspm_emu_asan.so")
dp = C.POINTER(C.c_double)
P = lambda a: a.ctypes.data_as(dp)
L.emu_corr3d_dist_f64.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double] + [C.c_int] * 5 + [dp, C.c_int, C.c_int]
L.emu_corr3d_dist_pruned_f64.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double] + [C.c_int] * 7 + [dp, C.c_int, C.c_int]
L.emu_corr3d_f64.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double] + [C.c_int] * 5 + [dp, C.c_int]
L.emu_corr3d_pruned_f64.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double] + [C.c_int] * 7 + [dp, C.c_int]
rng = np.random.default_rng(0)
n = 0
for shape, win, G in [((4, 4, 4), (0, 4), 1), ((6, 8, 8), (3, 7), 2), ((8, 16, 12), (2, 9), 4), ((12, 24, 10), (0, 5), 3),
                      ((16, 16, 6), (5, 6), 8), ((10, 20, 28), (7, 21), 5), ((14, 28, 16), (0, 16), 2), ((9, 12, 8), (1, 8), 3)]:
    a = np.ascontiguousarray(rng.random(shape)); b = np.ascontiguousarray(rng.random(shape) + 0.1)
    out = np.empty(shape[:2] + (win[1] - win[0],))
    for ts in (-1, 0, 1, 2, 1000, 1001, 1003):
        rc = L.emu_corr3d_dist_f64(P(a), P(b), 1.08, 1.08, 1.0, *shape, win[0], win[1], P(out), G, ts); assert rc == 0, rc; n += 1
for shape, win, sup, G in [((8, 8, 64), (20, 30), (60, 9), 2), ((6, 12, 48), (10, 22), (44, 8), 3), ((8, 16, 40), (3, 9), (0, 5), 4),
                           ((4, 8, 32), (0, 32), (30, 4), 2), ((12, 8, 96), (40, 61), (90, 13), 2)]:
    a = np.ascontiguousarray(rng.random(shape)); b = np.zeros(shape); b[:, :, (sup[0] + np.arange(sup[1])) % shape[2]] = 0.5
    out = np.empty(shape[:2] + (win[1] - win[0],))
    for ts in (-1, 0, 2, 1000, 1002):
        rc = L.emu_corr3d_dist_pruned_f64(P(a), P(b), 1.08, 1.08, 1.0, *shape, win[0], win[1], sup[0], sup[1], P(out), G, ts); assert rc > 0, rc; n += 1
        rc = L.emu_corr3d_pruned_f64(P(a), P(b), 1.08, 1.08, 1.0, *shape, win[0], win[1], sup[0], sup[1], P(out), ts if ts < 1000 else -1); assert rc > 0, rc; n += 1
for shape, win in [((4, 4, 4), (0, 4)), ((7, 9, 10), (1, 2)), ((12, 10, 28), (0, 1)), ((14, 15, 16), (5, 12)), ((16, 3, 6), (0, 5)), ((6, 5, 9), (2, 7))]:
    a = np.ascontiguousarray(rng.random(shape)); b = np.ascontiguousarray(rng.random(shape))
    out = np.empty(shape[:2] + (win[1] - win[0],))
    for ts in (-1, 0, 1, 2, 100, 101, 103, 200, 201, 203):
        rc = L.emu_corr3d_f64(P(a), P(b), 1.08, 1.0, 1.0, *shape, win[0], win[1], P(out), ts); assert rc == 0, rc; n += 1
print("asan clean:", n, "runs")
