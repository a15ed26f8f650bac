import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from dbspm_b200 import synth, interactions as ops
from dbspm_b200.sweep import fdbm_sweep
cfg = synth.CONFIGS["tiny"]
rhoS, potS, atoms = synth.make_sample(cfg); rhoT, nrhoT = synth.make_tip(cfg, variant="noisy")
dr = np.array(cfg.cell) / np.array(cfg.shape)
sets = [(1.0, 0.4, 0.2), (1.08, 0.4, 0.2), (1.08, 0.6, 0.24), (1.0, 0.6, 0.24)]
dev = torch.device("cuda", 0)
out = fdbm_sweep(rhoS.to(dev), potS.to(dev), rhoT.to(dev), nrhoT.to(dev), dr, (24, 32), 6, sets, cfg.rpivot, keep_static=True)
torch.cuda.synchronize()
for (alpha, V0, kappa), r in zip(sets, out["relax"]):
    static = np.zeros((28, 30, 8)); static += out["sr"][alpha].cpu().numpy() * V0; static += out["es"].cpu().numpy()
    print("static equal", np.array_equal(static, r["static"].cpu().numpy()), np.abs(static - r["static"].cpu().numpy()).max())
    ref = O.relax_tile(static, 0, 28, 0, 30, 6, kappa, cfg.rpivot, dr)
    r2 = ops.relax_grid(r["static"], kappa, cfg.rpivot, dr, 6)
    dE = np.abs(r["E"].cpu().numpy() - ref[..., 0]) / np.abs(ref[..., 0]).max()
    dE2 = np.abs(r2["E"].cpu().numpy() - ref[..., 0]) / np.abs(ref[..., 0]).max()
    print(alpha, V0, kappa, "sweep maxdE", dE.max(), (dE > 1e-6).mean(), " rerun maxdE", dE2.max())
