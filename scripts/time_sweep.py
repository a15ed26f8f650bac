"""Developer aid: the 16-set FDBM sweep of config 4 (pyridine grid), relax runs on 1 vs several streams."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dbspm_b200 import synth, sweep
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "pyridine"]
dr = np.array(cfg.cell) / np.array(cfg.shape)
rhoS, potS, atoms = synth.make_sample(cfg, device=dev)
rhoT, nrhoT = synth.make_tip(cfg, device=dev, variant="noisy")
vdw = synth.make_vdw_window(cfg, atoms, cfg.shape[:2] + (54,), 0.0, dev)
sets = [(a, v, 0.20 if i < 8 else 0.24) for i, (a, v) in enumerate((a, v) for a in (1.00, 1.04, 1.08, 1.12) for v in (35.0, 42.9071, 50.0, 60.0))]
for ns in (1, 2, 4, 8, 1, 4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = sweep.fdbm_sweep(rhoS, potS, rhoT, nrhoT, dr, (76, 130), 24, sets, cfg.rpivot, vdw=vdw, streams=ns)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    npos = cfg.shape[0] * cfg.shape[1] * 24 * len(sets)
    print(f"{cfg.name if hasattr(cfg, 'name') else ''} 16 sets, {ns} stream(s): {1e3 * (t1 - t0):.1f} ms, {npos / (t1 - t0) / 1e6:.1f} M positions/s, E[0]={float(r['relax'][0]['E'][3, 4, 5]):.12g}", flush=True)
