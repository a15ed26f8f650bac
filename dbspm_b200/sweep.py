"""Batched FDBM parameter sweep (BASELINE.json configs[3]; motivated by scripts/fdbmfit:136-154,
which re-runs the FFT correlation inside a least-squares loop).

For a list of (alpha, V0, kappa) sets: the es grid is computed once, one sr grid per distinct
alpha (both densities are raised to alpha, dbspm:329), and one relax per set
(static = V0*sr_alpha + es + vdw, dbspm:479-486).  All arrays stay in HBM between steps."""
from __future__ import annotations

import torch

from . import interactions as ops
from . import parallel


class SweepGraph:
    """The whole sweep captured in two CUDA graphs (correlations | static + relax of every set) and replayed.

    One pyridine-sized correlation is five launches of ~0.1 ms and a relax of 38 416 columns fills a third of the GPU: the
    sweep is launch- and host-bound (about 120 kernel launches through ctypes + torch allocations per sweep).  Capturing it
    once removes the host from the loop; the relax runs of the sets keep their fork/join over `streams` side streams inside
    the graph, so independent sets still overlap.  Inputs are read from the tensors given here (refresh them in place with
    .copy_() between replays); parameters (alpha, V0, kappa) are baked into the captured launches.

        plan = SweepGraph(rhoS, potS, rhoT, nrhoT, dr, (76, 130), 24, sets, 3.02, vdw=vdw)
        out = plan.run()          # {"es", "sr", "relax"}: the same tensors every replay
    """

    def __init__(self, rhoS, potS, rhoT, nrhoT, dr, z_window, nz_relax, param_sets, rpivot, vdw=None, dR=None, streams=4):
        self.args = dict(dr=dr, z_window=z_window, nz_relax=nz_relax, param_sets=list(param_sets), rpivot=rpivot, dR=dR, streams=streams)
        dev = rhoS.device
        self.inputs = tuple(ops._to_dev(a, dev) for a in (rhoS, potS, rhoT, nrhoT))
        self.vdw = ops._to_dev(vdw, dev) if vdw is not None else None
        self.dev = dev
        # eager warm-up on a side stream: FFT plans, function attributes and the allocator's pools exist before the capture
        warm = torch.cuda.Stream(device=dev)
        warm.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(warm):
            _corr_part(*self.inputs, dr, z_window, self.args["param_sets"])
            fdbm_sweep(*self.inputs, vdw=self.vdw, keep_static=False, _sync=False, _local=True, **self.args)
        torch.cuda.current_stream(dev).wait_stream(warm)
        torch.cuda.synchronize(dev)
        before = set(ops._WS_CACHE)
        self.g_corr, self.g_relax = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_corr):
            self.es, self.sr = _corr_part(*self.inputs, dr, z_window, self.args["param_sets"])
        with torch.cuda.graph(self.g_relax, pool=self.g_corr.pool()):
            self.relax = _relax_part(self.es, self.sr, self.vdw, dr, nz_relax, self.args["param_sets"], rpivot, dR, streams, False, dev)
        # workspaces allocated during capture live in the graphs' pool: keep them here, out of the eager cache
        self._ws = [ops._WS_CACHE.pop(k) for k in list(ops._WS_CACHE) if k not in before]

    def run_corr(self):
        self.g_corr.replay()

    def run_relax(self):
        self.g_relax.replay()

    def run(self):
        self.g_corr.replay()
        self.g_relax.replay()
        return {"es": self.es, "sr": self.sr, "relax": self.relax}


class ShardedSweepGraph:
    """The sweep under torchrun (one process per GPU) with the host out of the loop: the sets are dealt over the ranks
    (parallel.sweep_assignment: contiguous blocks after sorting by alpha), every rank captures ITS sets as a SweepGraph (its own
    es -- 0.35 ms, cheaper than a broadcast inside the step -- the sr grids of its alphas, its relax runs on forked streams)
    plus the copies of the requested results into one stacked buffer, and a replay is followed by ONE gather of the stacked
    buffers to rank 0 (received straight into place when the ranks hold equally many sets).  Same bits as the single-GPU
    sweep (every kernel is deterministic).

        plan = ShardedSweepGraph(rhoS, potS, rhoT, nrhoT, dr, (76, 130), 24, sets, 3.02, vdw=vdw)
        res = plan.run()      # rank 0: list of dicts in the order of `sets` (views into one buffer, refreshed by every run)
    """

    def __init__(self, rhoS, potS, rhoT, nrhoT, dr, z_window, nz_relax, param_sets, rpivot, vdw=None, dR=None, streams=4,
                 keys=("E", "tip_x", "tip_y", "tip_z"), dst=0):
        import torch.distributed as dist
        self.rank, self.world = parallel._world()
        self.keys, self.dst = tuple(keys), int(dst)
        self.param_sets = [tuple(float(v) for v in p) for p in param_sets]
        self.assignment = parallel.sweep_assignment(self.param_sets, self.world)
        mine = [self.param_sets[i] for i in self.assignment[self.rank]]
        dev = rhoS.device
        nx, ny = int(rhoS.shape[0]), int(rhoS.shape[1])
        self.counts = [len(a) for a in self.assignment]
        big = max(self.counts)
        # (sets, keys, nx, ny, nz): padded to the largest share so that one plain gather moves everything
        self.stack = torch.zeros((big, len(self.keys), nx, ny, int(nz_relax)), dtype=torch.float64, device=dev)
        self.plan = None
        if mine:
            self.plan = SweepGraph(rhoS, potS, rhoT, nrhoT, dr, z_window, nz_relax, mine, rpivot, vdw=vdw, dR=dR, streams=streams)
            self.g_stack = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_stack, pool=self.plan.g_corr.pool()):
                for j, r in enumerate(self.plan.relax):
                    for q, k in enumerate(self.keys):
                        self.stack[j, q].copy_(r[k])
        self.recv = None
        if self.world > 1 and self.rank == self.dst:
            self.recv = torch.empty((self.world,) + tuple(self.stack.shape), dtype=torch.float64, device=dev)
        self._dist = dist

    def run(self):
        if self.plan is not None:
            self.plan.run()
            self.g_stack.replay()
        if self.world == 1:
            parts = [self.stack]
        else:
            self._dist.gather(self.stack, list(self.recv.unbind(0)) if self.recv is not None else None, dst=self.dst)
            if self.rank != self.dst:
                return None
            parts = list(self.recv.unbind(0))
        out = [None] * len(self.param_sets)
        for a, part in zip(self.assignment, parts):
            for j, i in enumerate(a):
                d = {k: part[j, q] for q, k in enumerate(self.keys)}
                d.update(alpha=self.param_sets[i][0], V=self.param_sets[i][1], kappa=self.param_sets[i][2])
                out[i] = d
        return out


def _corr_part(A, P, T, NT, dr, z_window, param_sets):
    z_lo, z_hi = z_window
    es = ops.density_overlap_window(P, NT, dr, z_lo, z_hi)
    sr = {}
    for alpha, _, _ in param_sets:
        a = float(alpha)
        if a not in sr:
            sr[a] = ops.density_overlap_window(A, T, dr, z_lo, z_hi, a, a)
    return es, sr


def _relax_part(es, sr, vd, dr, nz_relax, param_sets, rpivot, dR, streams, keep_static, dev):
    # The relax of a small scan grid does not fill the GPU (pyridine: 1200 one-warp blocks for 2368 resident slots)
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
    return out


def fdbm_sweep(rhoS, potS, rhoT, nrhoT, dr, z_window, nz_relax, param_sets, rpivot, vdw=None, dR=None,
               keep_static=False, streams=4, _sync=True, _local=False):
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
    if world > 1 and not _local:      # (_local: this rank's own list, no collectives -- the warm-up of a captured sweep)
        return _sweep_sharded(A, P, T, NT, dr, z_lo, z_hi, nz_relax, param_sets, rpivot, vdw, dR, dev, rank, world)
    es, sr = _corr_part(A, P, T, NT, dr, (z_lo, z_hi), param_sets)
    vd = ops._to_dev(vdw, dev) if vdw is not None else None
    out = _relax_part(es, sr, vd, dr, nz_relax, param_sets, rpivot, dR, streams, keep_static, dev)
    if _sync and min(int(streams), len(param_sets)) > 1:
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
    my_sets = [param_sets[i] for i in assignment[rank]]
    sr = {}
    for alpha, _, _ in my_sets:
        a = float(alpha)
        if a not in sr:
            sr[a] = ops.density_overlap_window(A, T, dr, z_lo, z_hi, a, a)
    # the rank's own sets overlap on side streams like the single-GPU sweep (a pyridine-sized relax fills half a GPU)
    mine = _relax_part(es, sr, vd, dr, nz_relax, my_sets, rpivot, dR, 4, False, dev) if my_sets else []
    if len(my_sets) > 1:
        torch.cuda.synchronize(dev)
    for r in mine:
        for key in ("alpha", "V", "kappa"):
            r.pop(key, None)
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
