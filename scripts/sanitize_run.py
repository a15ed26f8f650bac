"""Developer aid: small end-to-end run of every kernel (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, tools, synth
rng = np.random.default_rng(0)
for shape, win in [((14, 15, 16), (5, 12)), ((12, 10, 28), (3, 10)), ((22, 26, 6), (0, 6)), ((196, 6, 24), (4, 15)), ((40, 120, 12), (0, 12))]:
    a, b = rng.random(shape), rng.random(shape) - 0.4
    for flags in (0, 2, 4):
        ops.density_overlap_window(a, b, np.ones(3), win[0], win[1], 1.08 if flags == 0 else 1.0, 1.0, flags=flags)
static = synth.make_static_like((40, 36, 54), 5).numpy()
dr = np.array([0.0765306, 0.0765306, 0.075])
r = ops.relax_grid(static, 0.2, 3.02, dr, 6, want_nfev=True)
r2 = ops.relax_grid(static, 0.2, 3.02, dr, 6, coalesce=False, dR=np.diag(dr), tile=(3, 20, 5, 30))
tools.get_df_image(r["E"], dr[2], n=4)
ops.TricubicField(static).gather(rng.uniform(-5, 60, (100, 3)))
ops.build_static([(static, 2.0), (static, 1.0)])
torch.cuda.synchronize()
print("sanitize run ok")
