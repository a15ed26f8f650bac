import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import oracle as O
from dbspm_b200 import synth, interactions as ops
cfg = synth.CONFIGS["tiny"]
rhoS, potS, atoms = synth.make_sample(cfg); rhoT, nrhoT = synth.make_tip(cfg, variant="noisy")
dr = np.array(cfg.cell) / np.array(cfg.shape)
es = O.density_overlap_fft(potS.numpy(), nrhoT.numpy(), dr)[:, :, 24:32]
for alpha, V0, kappa in [(1.0, 0.4, 0.2), (1.08, 0.4, 0.2), (1.08, 0.6, 0.24), (1.0, 0.6, 0.24)]:
    sr = O.density_overlap_fft(rhoS.numpy() ** alpha, rhoT.numpy() ** alpha, dr)[:, :, 24:32]
    static = np.zeros(sr.shape); static += sr * V0; static += es
    ref, nf = O.relax_tile(static, 0, 28, 0, 30, 6, kappa, cfg.rpivot, dr, want_nfev=True)
    for co in (False, True):
        r = ops.relax_grid(static, kappa, cfg.rpivot, dr, 6, want_nfev=True, coalesce=co)
        dE = np.abs(r["E"] - ref[..., 0]) / np.abs(ref[..., 0]).max()
        want = np.stack(O.sph2cart(ref[..., 1], ref[..., 2], cfg.rpivot))
        got = np.stack([r["tip_x"], r["tip_y"], r["tip_z"]])
        dp = np.abs(got - want).max(axis=0) / cfg.rpivot
        bad = np.argwhere(dE > 1e-6)
        print(alpha, V0, kappa, co, "fracE", (dE > 1e-6).mean(), "maxdE", dE.max(), "fracpos", (dp > 1e-6).mean(), "nfev diff", np.abs(r["nfev"] - nf).max(), "bad idx", bad[:5].tolist())
