"""Batched FDBM parameter sweep (BASELINE.json configs[3]; motivated by scripts/fdbmfit:136-154,
which re-runs the FFT correlation inside a least-squares loop).

For a list of (alpha, V0, kappa) sets: the es grid is computed once, one sr grid per distinct
alpha (both densities are raised to alpha, dbspm:329), and one relax per set
(static = V0*sr_alpha + es + vdw, dbspm:479-486).  All arrays stay in HBM between steps."""
from __future__ import annotations

import torch

from . import interactions as ops
from . import parallel


def fdbm_sweep(rhoS, potS, rhoT, nrhoT, dr, z_window, nz_relax, param_sets, rpivot, vdw=None, dR=None,
               keep_static=False, streams=4):
    """param_sets: iterable of (alpha, V0, kappa).  Inputs: CUDA tensors (or numpy, uploaded once).
    Returns {"es": window, "sr": {alpha: window}, "relax": [dict per set]} with CUDA tensors.
    streams: how many CUDA streams the independent relax runs are spread over (1 = serial).

    Under torchrun (torch.distributed initialised, one process per GPU) the sets are dealt over the
    ranks (parallel.sweep_assignment); es is computed on rank 0 and broadcast, every rank computes the
    sr grids of its own alphas, and rank 0 returns all relax results in the order of `param_sets`
    (the other ranks return only their own sets, `None` elsewhere)."""
    dev = rhoS.device if isinstance(rhoS, torch.Tensor) and rhoS.is_cuda else ops._device()
    A, P, T, NT = (ops._to_dev(a, dev) for a in (rhoS, potS, rhoT, nrhoT))
    z_lo, z_hi = z_window
    param_sets = list(param_sets)
    rank, world = parallel._world()
    if world > 1:
        return _sweep_sharded(A, P, T, NT, dr, z_lo, z_hi, nz_relax, param_sets, rpivot, vdw, dR, dev, rank, world)
    es = ops.density_overlap_window(P, NT, dr, z_lo, z_hi)
    vd = ops._to_dev(vdw, dev) if vdw is not None else None
    sr = {}
    for alpha, _, _ in param_sets:
        a = float(alpha)
        if a not in sr:
            sr[a] = ops.density_overlap_window(A, T, dr, z_lo, z_hi, a, a)
    # The relax of a small scan grid does not fill the GPU (pyridine: 300 blocks for 592 resident slots)
    # and the sets are independent: run them on a few streams so that their kernels overlap.
    main = torch.cuda.current_stream(dev)
    nstreams = max(1, min(int(streams), len(param_sets)))
    side = [torch.cuda.Stream(device=dev) for _ in range(nstreams)] if nstreams > 1 else [main]
    out = []
    for i, (alpha, V0, kappa) in enumerate(param_sets):
        a = float(alpha)
        st = side[i % nstreams]
        st.wait_stream(main)
        with torch.cuda.stream(st):
            comps = [(sr[a], float(V0)), (es, 1.0)] + ([(vd, 1.0)] if vd is not None else [])
            static = ops.build_static(comps)
            r = ops.relax_grid(static, float(kappa), rpivot, dr, nz_relax, dR=dR)
        r.update(alpha=a, V=float(V0), kappa=float(kappa))
        if keep_static:
            r["static"] = static
        out.append(r)
    for st in side:
        main.wait_stream(st)
    if nstreams > 1:
        torch.cuda.synchronize(dev)     # results were allocated on side streams: make them safe for any consumer
    return {"es": es, "sr": sr, "relax": out}


def _sweep_sharded(A, P, T, NT, dr, z_lo, z_hi, nz_relax, param_sets, rpivot, vdw, dR, dev, rank, world):
    nx, ny, _ = (int(v) for v in A.shape)
    if rank == 0:
        es = ops.density_overlap_window(P, NT, dr, z_lo, z_hi)
    else:
        es = torch.empty((nx, ny, z_hi - z_lo), dtype=torch.float64, device=dev)
    parallel.broadcast_(es)
    vd = ops._to_dev(vdw, dev) if vdw is not None else None
    assignment = parallel.sweep_assignment(param_sets, world)
    sr, mine = {}, []
    for i in assignment[rank]:
        alpha, V0, kappa = param_sets[i]
        a = float(alpha)
        if a not in sr:
            sr[a] = ops.density_overlap_window(A, T, dr, z_lo, z_hi, a, a)
        comps = [(sr[a], float(V0)), (es, 1.0)] + ([(vd, 1.0)] if vd is not None else [])
        mine.append(ops.relax_grid(ops.build_static(comps), float(kappa), rpivot, dr, nz_relax, dR=dR))
    gathered = parallel.gather_sweep(mine, assignment)
    if rank == 0:
        for i, r in enumerate(gathered):
            r.update(alpha=float(param_sets[i][0]), V=float(param_sets[i][1]), kappa=float(param_sets[i][2]))
        relax = gathered
    else:
        relax = [None] * len(param_sets)
        for i, r in zip(assignment[rank], mine):
            r.update(alpha=float(param_sets[i][0]), V=float(param_sets[i][1]), kappa=float(param_sets[i][2]))
            relax[i] = r
    return {"es": es, "sr": sr, "relax": relax}
