"""Developer aid: time the relax kernel of one library build (DBSPM_B200_LIB selects it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import interactions as ops, synth, _lib
dev = torch.device("cuda", 0)
CO = os.environ.get('RELAX_COALESCE', '1') == '1'
for shape in [tuple(int(v) for v in os.environ.get("RELAX_SHAPE", "1024,1024,54").split(","))]:
    static = synth.make_static_like(shape, 3, dev)
    dr = np.array([0.075, 0.075, 0.075])
    ops.relax_grid(static, 0.2, 3.02, dr, 24, coalesce=CO)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.relax_grid(static, 0.2, 3.02, dr, 24, coalesce=CO); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = float(np.median(ts))
    print(f"{os.path.basename(_lib.LIB_PATH)} coalesce={CO} {shape}: relax {t:.2f} ms {shape[0]*shape[1]*24/t/1e3:.1f} Mpos/s", flush=True)
