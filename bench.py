#!/usr/bin/env python3
"""bench.py -- sr+es grid pts/s and tip positions relaxed/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One step = one pass of the hot path over the workload: sr (alpha-powered overlap), es,
static = V*sr + es + vdw, relax of every tip position.  Default workload = BASELINE.json
configs[4], the 1024x1024x256 slab (it fits one GPU, so N = 1, 2, 4, 8 all run it: strong
scaling).  N >= 2: ONE group of all ranks computes sr, then es, with the transform itself
distributed over x-slabs (parallel.sr_es_distributed: both exchanges fused into the passes as
peer stores over NVLink); every rank keeps its x-slab of both windows, assembles its static
tile, fetches a halo of the lever length from its two neighbours and relaxes its x-tile
(relax kernel in x-slab mode); the windows and the relax tiles go to rank 0 for the .npz, the
window gathers beside the relax kernel.  Once per run the assembled es window is checked
against the single-GPU path and the relaxed tile against the whole-window relax (same bits).
--legacy-groups: the round-1 plan (an sr and an es group, windows gathered and broadcast back).
Timed with CUDA events on the launching stream (torch's current stream), max over ranks.
The arrays (2.1 GB each) are far larger than the 126 MB L2, so no explicit flush is needed.

`value` = 2*Nx*Ny*Nz / (t_sr + t_es) with inputs resident in HBM; `relax.value` = positions/s;
`e2e` = the same metric with host (pinned) inputs copied in and results copied out inside the
timed region; `roofline` = algorithmic bytes 8*(2N + N_w) per interaction / measured time
against MEASURED_PEAKS.json; `cpu_baseline` = the oracle port on this box's host cores on a
bounded sample.  --impl reference prints the CPU arm as its own line (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sr+es grid pts/s"
UNIT = "grid pts/s"


# ------------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(workload):
    """DRAM bytes (read + write) of one interaction's five launches from the committed ncu
    capture (profiles/traffic.json); None when that workload has not been profiled."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return float(json.load(fh)[workload]["total_per_interaction"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
_G = {}


def _relax_worker(rows):
    from oracle import oracle as O
    i_lo, i_hi = rows
    return O.relax_tile(_G["static"], i_lo, i_hi, 0, _G["nj"], _G["nz"], _G["kappa"], _G["rpivot"], _G["dr"])[..., 0].sum()


def cpu_reference_pass(sample_shape=(256, 256, 128), relax_cols=(64, 64), seed=20260104):
    """One bounded pass of the reference algorithm on the host cores.

    sr/es: the numpy restatement of get_density_overlap (interactions.py:9-22), ONE process --
    the reference runs these steps on rank 0 only (dbspm:320,339).  relax: the C restatement of
    relax_tip_z + scipy Powell + tricubic over all host cores (the reference spreads lateral
    points over MPI ranks, dbspm:512-514; its own per-point cost is Python-bound, ~1 ms/position
    in the survey, so the C port is a *stronger* baseline than the reference itself)."""
    import multiprocessing as mp

    from dbspm_b200 import synth
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    nx, ny, nz = sample_shape
    a, b = rng.random(sample_shape), rng.random(sample_shape)
    dr = np.array([0.075, 0.075, 0.075])
    z_lo, z_hi = 38, 92
    t0 = time.perf_counter()
    sr = O.density_overlap_fft(a ** 1.08, b ** 1.08, dr)[:, :, z_lo:z_hi]
    t1 = time.perf_counter()
    es = O.density_overlap_fft(a, b - 0.5, dr)[:, :, z_lo:z_hi]
    t2 = time.perf_counter()
    ni, nj = relax_cols
    static = synth.make_static_like((ni, nj, 54), seed).numpy()
    cores = os.cpu_count() or 1
    _G.update(static=static, nj=nj, nz=24, kappa=0.2, rpivot=3.02, dr=dr)
    O.lib()
    chunks = [(g * ni // (4 * cores), (g + 1) * ni // (4 * cores)) for g in range(4 * cores)]
    chunks = [c for c in chunks if c[1] > c[0]]
    t3 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_relax_worker, chunks)
    t4 = time.perf_counter()
    npts = nx * ny * nz
    return {
        "t_sr": t1 - t0, "t_es": t2 - t1, "t_relax": t4 - t3, "grid_pts": npts, "positions": ni * nj * 24,
        "pts_per_s": 2 * npts / (t2 - t0), "pos_per_s": ni * nj * 24 / (t4 - t3), "cores_relax": cores,
        "sample": f"sr+es: numpy complex fftn on a {nx}x{ny}x{nz} f64 grid, 1 process (as the reference); "
                  f"relax: {ni}x{nj} columns x 24 heights, C port of scipy-Powell+tricubic on {cores} processes",
    }


def cpu_full_size_overlap(shape, z_lo, z_hi, alpha, seed=20260104):
    """sr and es ONCE at the benchmarked shape with the numpy restatement of get_density_overlap (interactions.py:9-22,
    complex fftn of the whole grid, index reversal, roll, slice), one process as the reference runs them (dbspm:320,339).
    Random positive densities: the cost of the transform does not depend on the values.  ~20 GB of temporaries."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    dr = np.array([0.075, 0.075, 0.075])
    a = rng.random(shape)
    b = rng.random(shape)
    t0 = time.perf_counter()
    sr = O.density_overlap_fft(a ** alpha, b ** alpha, dr)[:, :, z_lo:z_hi].copy()
    t1 = time.perf_counter()
    b -= 0.5
    es = O.density_overlap_fft(a, b, dr)[:, :, z_lo:z_hi].copy()
    t2 = time.perf_counter()
    return {"t_sr": t1 - t0, "t_es": t2 - t1, "grid_pts": int(np.prod(shape)), "checksum": float(sr.sum() + es.sum())}


def run_reference_arm(args):
    """CPU arm: the oracle port of the reference's own algorithm on this box's host cores (the reference is pure Python
    with no compiled part).  sr+es: ONE pass at the benchmarked grid (same config as the GPU arm);
    relax: a bounded sample of lateral columns on all cores; the bounded sr+es sample of round 1 stays as a secondary field."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from dbspm_b200 import synth
    cfg = synth.CONFIGS[args.workload]
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_pass(sample_shape=(64, 64, 64), relax_cols=(8, 8))
    small = cpu_reference_pass()
    full = None
    note = "bounded sample only"
    try:
        import psutil
        need = 12 * 8 * int(np.prod(cfg.shape))
        if args.reference_full and psutil.virtual_memory().available > 1.3 * need:
            full = cpu_full_size_overlap(cfg.shape, 76, 130, cfg.alpha)
            note = "one full-size pass"
    except Exception as exc:                                   # noqa: BLE001
        note = f"full-size pass failed ({exc!r}); bounded sample only"
    if full is not None:
        t_corr = full["t_sr"] + full["t_es"]
        value = 2 * full["grid_pts"] / t_corr
        sample = (f"sr+es: numpy complex fftn restatement of get_density_overlap on the full {cfg.shape[0]}x{cfg.shape[1]}x{cfg.shape[2]} "
                  f"f64 grid, ONE pass, 1 process (as the reference); relax: " + small["sample"].split("relax: ")[1])
    else:
        t_corr = small["t_sr"] + small["t_es"]
        value = small["pts_per_s"]
        sample = small["sample"]
    t_rel_full = cfg.shape[0] * cfg.shape[1] * 24 / small["pos_per_s"]     # relax of the whole grid, extrapolated from the sample
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (t_corr + t_rel_full), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "cpu_sample": sample, "same_config_as_gpu_arm": full is not None,
                   "passes": note, "steps_run": 1,
                   "ms_per_step_note": "sr+es measured" + (" at full size" if full is not None else " on the sample")
                                       + " + relax of the whole grid extrapolated linearly from the sampled columns"},
        "relax": {"value": small["pos_per_s"], "unit": "tip positions/s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "ms_sr": 1e3 * (full or small)["t_sr"], "ms_es": 1e3 * (full or small)["t_es"],
                         "bounded_sample_value": small["pts_per_s"], "bounded_sample": small["sample"],
                         "relax_value": small["pos_per_s"], "relax_unit": "tip positions/s", "relax_cores": small["cores_relax"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
SWEEP_SETS = [(a, v, 0.20 if i < 8 else 0.24)
              for i, (a, v) in enumerate((a, v) for a in (1.00, 1.04, 1.08, 1.12) for v in (35.0, 42.9071, 50.0, 60.0))]


def run_sweep_arm(args):
    """BASELINE.json configs[3]: 16 (alpha, V0, k) sets over the pyridine grid -- es once, one sr per distinct alpha (4), 16
    relax runs.  N = 1: the sweep captured in CUDA graphs (dbspm_b200.sweep.SweepGraph) and replayed; the eager form is timed
    beside it.  N > 1: the sets dealt over the ranks (sweep._sweep_sharded: es broadcast, gather of the relax results)."""
    import torch
    import torch.distributed as dist

    from dbspm_b200 import interactions as ops
    from dbspm_b200 import parallel, synth, sweep
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, world, dev = parallel.init_from_env()
    cfg = synth.CONFIGS["pyridine"]
    nx, ny, nz = cfg.shape
    N, z_lo, z_hi, nz_relax = nx * ny * nz, 76, 130, 24
    dr = np.array(cfg.cell) / np.array(cfg.shape)
    rhoS, potS, atoms = synth.make_sample(cfg, dev)
    rhoT, nrhoT = synth.make_tip(cfg, dev, "noisy")
    vdw = synth.make_vdw_window(cfg, atoms, (nx, ny, z_hi - z_lo), 2.7, dev)
    npos = nx * ny * nz_relax * len(SWEEP_SETS)

    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def eager():
        return sweep.fdbm_sweep(rhoS, potS, rhoT, nrhoT, dr, (z_lo, z_hi), nz_relax, SWEEP_SETS, cfg.rpivot, vdw=vdw)

    for _ in range(max(args.warmup, 1)):
        eager()
    barrier()
    ts = []
    for _ in range(args.steps):
        a = ev(); out = eager(); b = ev(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_eager = float(t.item())
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if world == 1:
        plan = sweep.SweepGraph(rhoS, potS, rhoT, nrhoT, dr, (z_lo, z_hi), nz_relax, SWEEP_SETS, cfg.rpivot, vdw=vdw)
        for _ in range(max(args.warmup, 3)):
            plan.run()
        torch.cuda.synchronize()
        # same bits as the eager sweep
        same = all(torch.equal(plan.relax[i]["E"], out["relax"][i]["E"]) for i in range(len(SWEEP_SETS))) and torch.equal(plan.es, out["es"])
        tc, tr = [], []
        with ClockSampler(dev.index) as clk:
            for _ in range(args.steps):
                a = ev(); plan.run_corr(); b = ev(); plan.run_relax(); c = ev(); torch.cuda.synchronize()
                tc.append(a.elapsed_time(b)); tr.append(b.elapsed_time(c))
        t_corr, t_relax = float(np.mean(tc)), float(np.mean(tr))
        # end to end: pinned host inputs in, every set's E and tip positions out, inside the timed region
        hosts = [torch.empty(x.shape, dtype=torch.float64).pin_memory().copy_(x) for x in (rhoS, potS, rhoT, nrhoT)]
        outs_h = torch.empty((len(SWEEP_SETS), 4, nx, ny, nz_relax), dtype=torch.float64).pin_memory()
        te = []
        for it in range(args.steps + 1):
            a = ev()
            for d, h in zip(plan.inputs, hosts):
                d.copy_(h, non_blocking=True)
            plan.run()
            for i, r in enumerate(plan.relax):
                for q, key in enumerate(("E", "tip_x", "tip_y", "tip_z")):
                    outs_h[i, q].copy_(r[key], non_blocking=True)
            b = ev(); torch.cuda.synchronize()
            if it:
                te.append(a.elapsed_time(b))
        peak, peak_src = measured_peak()
        balg = 8 * (2 * N + nx * ny * (z_hi - z_lo))
        line.update(value=5 * N / (t_corr * 1e-3), ms_per_step=t_corr + t_relax,
                    relax={"value": npos / (t_relax * 1e-3), "unit": "tip positions/s", "ms": t_relax, "sets": len(SWEEP_SETS)},
                    sweep={"ms_correlations_es_plus_4_sr": t_corr, "ms_16_relax": t_relax, "ms_eager_whole_sweep": t_eager,
                           "graph_equals_eager_bit_for_bit": bool(same),
                           "note": "two CUDA graphs (5 correlations = 25 launches | 16 x (3 accumulate, re-layout, relax, sph2cart) = 96 "
                                   "launches on 4 forked streams), replayed; eager = the same calls issued from Python"},
                    roofline={"bound": "hbm", "achieved": 5 * balg / (t_corr * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": 5 * balg / (t_corr * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                              "note": "five pyridine-sized correlations (164 MB algorithmic each) are launch-latency sized: 25 launches of ~0.1 ms"},
                    e2e={"value": 5 * N / (float(np.mean(te)) * 1e-3), "unit": UNIT, "ms_whole_sweep": float(np.mean(te)),
                         "h2d_bytes_per_step": 4 * 8 * N, "d2h_bytes_per_step": int(outs_h.numel()) * 8},
                    gpu_launches=(25 + 96) * args.steps, clocks=clk.summary(),
                    config={"workload": "fdbmfit-style batch sweep (BASELINE.json configs[3]): 16 (alpha, V0, k) sets over the synthetic "
                                        "pyridine grid 196x196x240, window 54 planes, relax 196x196x24 per set", "sets": SWEEP_SETS})
    else:
        # every rank replays the CUDA graph of its own sets, then one gather of the stacked results to rank 0
        plan = sweep.ShardedSweepGraph(rhoS, potS, rhoT, nrhoT, dr, (z_lo, z_hi), nz_relax, SWEEP_SETS, cfg.rpivot, vdw=vdw)
        for _ in range(max(args.warmup, 3)):
            res = plan.run()
        barrier()
        same = None
        if rank == 0:      # same bits as the eager sharded sweep (whose results rank 0 holds in `out`)
            same = all(torch.equal(res[i]["E"], out["relax"][i]["E"]) and torch.equal(res[i]["tip_z"], out["relax"][i]["tip_z"])
                       for i in range(len(SWEEP_SETS)))
        tg = []
        with ClockSampler(dev.index) as clk:
            for _ in range(args.steps):
                barrier(); a = ev(); plan.run(); b = ev(); torch.cuda.synchronize(); tg.append(a.elapsed_time(b))
        t = torch.tensor([float(np.mean(tg))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_graph = float(t.item())
        n_mine = len(plan.assignment[rank])
        n_alpha = len({SWEEP_SETS[i][0] for i in plan.assignment[rank]})
        line.update(value=None, ms_per_step=t_graph,
                    relax={"value": npos / (t_graph * 1e-3), "unit": "tip positions/s (whole sweep time, correlations and gather included)"},
                    sweep={"ms_graph_whole_sweep": t_graph, "ms_eager_whole_sweep": t_eager, "graph_equals_eager_bit_for_bit": same,
                           "assignment": plan.assignment,
                           "note": "sets dealt over the ranks by alpha; every rank replays the CUDA graphs of its own sets (its own es, the "
                                   "sr grids of its alphas, its relax runs on 4 forked streams, copies into one stacked buffer), then ONE "
                                   "gather of the stacked results to rank 0; eager = sweep.fdbm_sweep issued from Python (es on rank 0 + "
                                   "broadcast, pickled shape exchange, padded gather)"},
                    gpu_launches=((1 + n_alpha) * 5 + 6 * n_mine + 4 * n_mine) * args.steps, clocks=clk.summary(),
                    config={"workload": "fdbmfit-style batch sweep (BASELINE.json configs[3]): 16 (alpha, V0, k) sets over the synthetic "
                                        "pyridine grid 196x196x240, sharded over %d ranks" % world, "sets": SWEEP_SETS})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


WORKLOAD_DESC = {
    "large_slab": "synthetic large slab 1024x1024x256 (BASELINE.json configs[4]), window 54 planes, relax 1024x1024x24",
    "tma_cu111": "synthetic 384x384x160 (BASELINE.json configs[2]), window 54 planes, relax 384x384x24",
    "pyridine": "synthetic pyridine-sized 196x196x240 (BASELINE.json configs[0] shape), window 54, relax 196x196x24",
    "triangulene": "synthetic triangulene-sized 240x320x240 (BASELINE.json configs[1] shape)",
    "sweep16": "fdbmfit-style batch sweep: 16 (alpha, V0, k) sets over the pyridine grid (BASELINE.json configs[3])",
}
SKIP = {"x_fwd": 1 << 8, "y_fwd": 1 << 9, "z_fused": 1 << 10, "y_inv": 1 << 11, "x_inv": 1 << 12}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from dbspm_b200 import interactions as ops
    from dbspm_b200 import parallel, synth
    from dbspm_b200.grid import Grid, get_grid_from_params
    from dbspm_b200.input_parser import default_params
    from argparse import Namespace

    # NCCL_DEBUG is left as the caller set it (the driver reads the rank count from NCCL's INFO lines); its log goes to
    # stderr unless the caller chose a file, so that stdout stays the one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank, world, dev = parallel.init_from_env()
    assert dev.type == "cuda", "bench.py needs a GPU (use --impl reference for the CPU arm)"
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    cfg = synth.CONFIGS[args.workload]
    nx, ny, nz = cfg.shape
    N = nx * ny * nz
    cell = np.diag(cfg.cell)
    gS = Grid(cfg.shape, cell=cell, positions=np.zeros((1, 3)), numbers=np.array([1]))
    p = Namespace(**{**default_params, "Z0": cfg.z0, "ZF": cfg.zf, "ZREF": cfg.zref, "RPIVOT": cfg.rpivot,
                     "ZPAD": cfg.zpad})
    win, nidx = get_grid_from_params(p, gS, nidx=True, rpivot=cfg.rpivot, zpad=cfg.zpad)
    _, _, nbox = get_grid_from_params(p, gS, nidx=True, nbox=True)
    z_lo, z_hi = int(nidx[2].start), int(nidx[2].stop)
    nzw = z_hi - z_lo
    nz_relax = int(nbox[1, 2] - nbox[0, 2])
    dr = gS.dr
    Nw = nx * ny * nzw

    # ---- synthetic inputs, generated on the device (identical on every rank) ---------------
    rhoS, potS, atoms = synth.make_sample(cfg, dev)
    rhoT, nrhoT = synth.make_tip(cfg, dev, "noisy")
    vdw = synth.make_vdw_window(cfg, atoms, (nx, ny, nzw), float(win.origin[2]), dev)
    torch.cuda.synchronize()

    # sr and es are independent: with N >= 2 half of the ranks compute sr and the other half es
    # (each half splits the window's planes); with N = 1 the single rank runs both
    plan = parallel.interaction_plan(world, z_lo, z_hi)
    jobs = plan[rank]
    # N >= 4: the transform itself is distributed inside each group (x-slabs, one all-to-all, gather of the packed
    # window rows); N = 2 has one rank per interaction and N = 1 runs both, so they keep the single-GPU five-pass path
    # N >= 2 (round 2): ONE group of all N ranks computes sr, then es, with the transform distributed over x-slabs; every rank
    # ends with the same x-slab of both windows, assembles its static tile, fetches a halo of the lever length from its two
    # neighbours and relaxes its tile -- the windows travel to rank 0 only for the .npz, beside the relax kernel.
    slab_mode = (world > 1 and not args.no_dist and not args.legacy_groups and nx % world == 0
                 and ops.dist_supported(cfg.shape, world))
    groups = None
    if slab_mode:
        groups = [list(range(world))]
    elif not args.no_dist:
        groups = parallel.dist_groups(world)
        if groups is not None and not ops.dist_supported(cfg.shape, len(groups[0])):
            groups = None
    if groups is not None:
        G = len(groups[0])
        my_which, my_h = rank // G, rank % G
        sx0, sx1 = parallel.x_slab(nx, G, my_h)
    halo = ops.relax_slab_halo(cfg.rpivot, dr) if slab_mode else 0
    asm, asm_side = None, None
    if slab_mode and not args.no_peer:
        shapes = {"sr": (nx, ny, nzw), "es": (nx, ny, nzw), "E": (nx, ny, nz_relax), "tip_x": (nx, ny, nz_relax),
                  "tip_y": (nx, ny, nz_relax), "tip_z": (nx, ny, nz_relax)}
        asm = {k: parallel.PeerAssembler.get(k, shp, dev) for k, shp in shapes.items()}
        if any(v is None for v in asm.values()):
            asm = None
        else:
            asm_side = torch.cuda.Stream(device=dev)
    dist_marks = []
    rest_marks = []        # slab mode: events inside "the rest" of a step (static | halo exchange | relax | tile gathers | window gathers)
    tiles = parallel.split_range(0, nx, world)
    my_il, my_ih = tiles[rank]
    inputs = {0: (rhoS, rhoT, cfg.alpha), 1: (potS, nrhoT, 1.0)}
    bufs = [] if slab_mode else [torch.empty((nx, ny, zh - zl), dtype=torch.float64, device=dev) for (_, zl, zh) in jobs]
    launches = {"n": 0}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def corr_phase(src, marks=None, supports=None):
        if groups is not None and supports is None:
            a, b, al = src[my_which]
            rec = [("start", ev())] if marks is not None else None
            wins = parallel.sr_es_distributed(ops, a[sx0:sx1], b[sx0:sx1], al, cfg.shape, dr, z_lo, z_hi, groups,
                                              on_stage=(lambda name: rec.append((name, ev()))) if rec is not None else None,
                                              peer=not args.no_peer)
            launches["n"] += 5                                     # y fwd | x fwd, z, x inv | y inv of this rank's x-slab
            if marks is not None:
                marks.append(ev())
                dist_marks.append(rec)
            return wins
        for (which, zl, zh), buf in zip(jobs, bufs):
            a, b, al = src[which]
            if zh > zl:
                ops.density_overlap_window(a, b, dr, zl, zh, al, al, out=buf,
                                           tip_support=supports[which] if supports else None)
                launches["n"] += 5 + (1 if (zh - zl) & 1 else 0)
            if marks is not None:
                marks.append(ev())
        return parallel.gather_windows(plan, bufs, nx, ny, z_lo, z_hi, dev)        # rank 0: [sr, es]

    def slab_interaction(which, rec=None, src=None, supports=None, sliced=False):
        a, b, al = (src or inputs)[which]
        if not sliced:
            a, b = a[sx0:sx1], b[sx0:sx1]
        return parallel.sr_es_distributed(ops, a, b, al, cfg.shape, dr, z_lo, z_hi, groups, gather=False,
                                          on_stage=(lambda name: rec.append((name, ev()))) if rec is not None else None,
                                          peer=not args.no_peer, tip_support=supports[which] if supports else None)

    def slab_step(timed, src=None, vdw_slab=None, sliced=False):
        """sr, es (x-slabs of all ranks), static tile + halo, relax of the tile; the windows go to rank 0 beside the relax"""
        e0 = ev()
        rec = [("start", e0)]
        sr_slab = slab_interaction(0, rec, src, sliced=sliced)
        e1 = ev()
        rec.append(("start", e1))
        es_slab = slab_interaction(1, rec, src, sliced=sliced)
        e2 = ev()
        launches["n"] += 10                                        # per interaction: y fwd | x fwd, z, x inv | y inv
        # The windows and the relax tiles go to rank 0 only for the .npz.  With symmetric memory every rank copies its rows
        # straight into rank 0's buffers (parallel.PeerAssembler: copy engines over NVLink, no SM), the windows on a side
        # stream beside static / halo exchange / relax; otherwise NCCL gathers (the window gathers asynchronously).
        use_asm = asm is not None
        if use_asm:
            cur = torch.cuda.current_stream(dev)
            asm_side.wait_stream(cur)
            for key, slab in (("sr", sr_slab), ("es", es_slab)):
                asm[key].begin(asm_side)
                asm[key].put(slab, asm_side)
        else:
            w_sr, work_sr = parallel.gather_x_slabs(sr_slab, nx, async_op=True)
            w_es, work_es = parallel.gather_x_slabs(es_slab, nx, async_op=True)
        static_slab = ops.build_static([(sr_slab, cfg.V), (es_slab, 1.0), (vdw[sx0:sx1] if vdw_slab is None else vdw_slab, 1.0)])
        ea = ev()
        static_loc = parallel.exchange_halo(static_slab, halo)
        eb = ev()
        r = ops.relax_grid(static_loc, cfg.kappa, cfg.rpivot, dr, nz_relax, tile=(sx0, sx1, 0, ny), x_slab=((sx0 - halo) % nx, nx))
        ec = ev()
        launches["n"] += 6                                         # 3 x accumulate, y-quad re-layout, relax, sph2cart
        if use_asm:
            for key in ("E", "tip_x", "tip_y", "tip_z"):
                asm[key].begin()
                asm[key].put(r[key])
            res = {key: asm[key].end() for key in ("E", "tip_x", "tip_y", "tip_z")}
            res = res if rank == 0 else None
            ed = ev()
            cur.wait_stream(asm_side)
            w_sr, w_es = asm["sr"].end(), asm["es"].end()
        else:
            res = parallel.relax_sharded(lambda il, ih: {k: r[k] for k in ("E", "tip_x", "tip_y", "tip_z")}, nx, dev)
            ed = ev()
            work_sr.wait(); work_es.wait()
        e3 = ev()
        if timed is not None:
            timed.append((e0, e1, e2, e3))
            dist_marks.append(rec)
            rest_marks.append((e2, ea, eb, ec, ed, e3))
        return {"static": static_loc, "w_sr": w_sr, "w_es": w_es, "res": res, "r": r, "ev": (e0, e2, e3)}

    def step(timed):
        if slab_mode:
            return slab_step(timed)["static"]
        e0 = ev()
        mid = []
        wins = corr_phase(inputs, mid)
        e2 = ev()
        if world > 1:                       # every rank needs the whole static window
            if rank != 0:
                wins = [torch.empty((nx, ny, nzw), dtype=torch.float64, device=dev) for _ in range(2)]
            parallel.broadcast_(wins[0]); parallel.broadcast_(wins[1])
        sr, es = wins
        static = ops.build_static([(sr, cfg.V), (es, 1.0), (vdw, 1.0)])
        launches["n"] += 3
        res = parallel.relax_sharded(
            lambda il, ih: {k: v for k, v in ops.relax_grid(static, cfg.kappa, cfg.rpivot, dr, nz_relax,
                                                            tile=(il, ih, 0, ny)).items() if k in ("E", "tip_x", "tip_y", "tip_z")},
            nx, dev)
        launches["n"] += 3 if my_ih > my_il else 0          # transpose + relax + cart
        e3 = ev()
        if timed is not None:
            timed.append((e0, mid[0], e2, e3))
        return static

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    for _ in range(args.warmup):
        step(None)
    barrier()
    # distributed transform: rank 0 checks the assembled es window against the single-GPU five-pass path once
    dist_check = None
    relax_check = None
    if slab_mode:
        # once per run: the es window assembled on rank 0 against the single-GPU five-pass path, and this rank's relaxed
        # energies (tile + halo form) against the same tile relaxed on the whole static window (must be the same bits)
        keep = slab_step(None)
        if rank == 0:
            one = ops.density_overlap_window(potS, nrhoT, dr, z_lo, z_hi)
            dist_check = float((keep["w_es"] - one).abs().max() / one.abs().max())
            sr1 = ops.density_overlap_window(rhoS, rhoT, dr, z_lo, z_hi, cfg.alpha, cfg.alpha)
            full_static = ops.build_static([(keep["w_sr"], cfg.V), (keep["w_es"], 1.0), (vdw, 1.0)])
            sub = ops.relax_grid(full_static, cfg.kappa, cfg.rpivot, dr, nz_relax, tile=(sx0, sx0 + 8, 0, ny))
            relax_check = bool(torch.equal(sub["E"], keep["r"]["E"][:8]) and torch.equal(sub["tip_x"], keep["r"]["tip_x"][:8]))
            # ... and what arrived on rank 0 from the LAST rank's tile (assembled through peer copies or the NCCL gather)
            sub2 = ops.relax_grid(full_static, cfg.kappa, cfg.rpivot, dr, nz_relax, tile=(nx - 8, nx, 0, ny))
            relax_check = relax_check and bool(torch.equal(sub2["E"], keep["res"]["E"][nx - 8:])
                                               and torch.equal(sub2["tip_z"], keep["res"]["tip_z"][nx - 8:])
                                               and torch.equal(keep["res"]["tip_y"][:8], keep["r"]["tip_y"][:8]))
            del sub2
            dist_check_sr = float((keep["w_sr"] - sr1).abs().max() / sr1.abs().max())
            dist_check = max(dist_check, dist_check_sr)
            del one, sr1, full_static, sub
        keep.clear()
        ops.release_workspace(); torch.cuda.empty_cache()
        barrier()
    elif groups is not None:
        wins = corr_phase(inputs, None)
        if rank == 0:
            one = ops.density_overlap_window(potS, nrhoT, dr, z_lo, z_hi)
            dist_check = float((wins[1] - one).abs().max() / one.abs().max())
            del one
        del wins
        barrier()
    # the checks above released the caches (they hold full-size copies): one more untimed step so that the allocator's pools
    # and the workspaces exist again before the timed region (a cudaMalloc inside a 16 ms step is not what is being measured)
    if world > 1:
        step(None)
        barrier()
    launches["n"] = 0
    dist_marks.clear()
    rest_marks.clear()
    marks = []
    with ClockSampler(dev.index) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            static = step(marks)
        barrier()
        t_wall = time.perf_counter() - t_wall0
    # N = 1: first job = sr, then es, no gather.  N > 1: sr and es run concurrently on different ranks;
    # only their sum (start -> both windows assembled on rank 0) is meaningful
    t_sr = sum(a.elapsed_time(b) for a, b, _, _ in marks) / args.steps
    t_es = sum(b.elapsed_time(c) for _, b, c, _ in marks) / args.steps
    t_rest = sum(c.elapsed_time(d) for _, _, c, d in marks) / args.steps
    t_all = sum(a.elapsed_time(d) for a, _, _, d in marks) / args.steps
    t_sr, t_es, t_rest, t_all = max_over_ranks([t_sr, t_es, t_rest, t_all])
    n_launch = launches["n"]
    rest_stages = None
    if rest_marks:
        names = ("static", "halo_exchange", "relax", "tile_gathers", "window_gathers_wait")
        mine = [sum(m[i].elapsed_time(m[i + 1]) for m in rest_marks[-args.steps:]) / args.steps for i in range(5)]
        rest_stages = dict(zip(names, max_over_ranks(mine)))
        print(f"rank {rank} rest of the step (ms): " + ", ".join(f"{n} {v:.2f}" for n, v in zip(names, mine)), file=sys.stderr, flush=True)

    # ---- per-launch durations of the correlation (stage-skip flags), this rank's first job ----
    stages = {}
    relax_ms = None
    which0, zl0, zh0 = jobs[0]
    if zh0 > zl0 and not slab_mode:
        allskip = sum(SKIP.values())
        for name, bit in SKIP.items():
            ts = []
            for it in range(4):
                a = ev()
                ops.density_overlap_window(potS, nrhoT, dr, zl0, zh0, out=bufs[0], flags=allskip & ~bit)
                b = ev(); torch.cuda.synchronize()
                if it:
                    ts.append(a.elapsed_time(b))
            stages[name] = float(np.mean(ts))
        ts = []
        for it in range(3):
            a = ev()
            ops.density_overlap_window(rhoS, rhoT, dr, zl0, zh0, cfg.alpha, cfg.alpha, out=bufs[0],
                                       flags=allskip & ~SKIP["x_fwd"])
            b = ev(); torch.cuda.synchronize()
            if it:
                ts.append(a.elapsed_time(b))
        stages["x_fwd_pow"] = float(np.mean(ts))
    ts = []
    for it in range(3):
        a = ev()
        r = ops.relax_grid(static, cfg.kappa, cfg.rpivot, dr, nz_relax, tile=(my_il, my_ih, 0, ny), want_nfev=(it == 0),
                           x_slab=((sx0 - halo) % nx, nx) if slab_mode else None)
        b = ev(); torch.cuda.synchronize()
        if it == 0:
            nfev_mean = float(r["nfev"].float().mean()) if r["nfev"].numel() else 0.0
        else:
            ts.append(a.elapsed_time(b))
    relax_ms = float(np.mean(ts))
    relax_ms_max, = max_over_ranks([relax_ms])

    # ---- the same sr+es with the compact tip variant (exactly zero beyond 3.5 A): z-pruned path ----
    compact = None
    if not args.no_compact:
        rhoTc, nrhoTc = synth.make_tip(cfg, dev, "compact")
        sup = {0: ops.tip_zsupport(rhoTc), 1: ops.tip_zsupport(nrhoTc)}
        src_c = {0: (rhoS, rhoTc, cfg.alpha), 1: (potS, nrhoTc, 1.0)}
        def compact_pass():
            if slab_mode:          # the distributed transform with the z-range cut (dbspm_corr3d_dist_fwd_peer_cut_f64)
                slab_interaction(0, None, src_c, sup)
                slab_interaction(1, None, src_c, sup)
            else:
                corr_phase(src_c, None, sup)
        for _ in range(2):
            compact_pass()
        barrier()
        ts = []
        for _ in range(args.steps):
            a = ev()
            compact_pass()
            b = ev(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        barrier()
        t_c, = max_over_ranks([float(np.mean(ts))])
        compact = {"value": 2 * N / (t_c * 1e-3), "unit": UNIT, "ms_sr_plus_es": t_c,
                   "tip_support_planes": {"rhoT": list(sup[0]), "nrhoT": list(sup[1])},
                   "note": "tip arrays exactly zero outside the listed (start, length) z-planes: only the planes of the "
                           "sample the window can see are transformed (dbspm_corr3d_pruned_f64); support scan not timed "
                           "(once per tip)"}
        del rhoTc, nrhoTc, src_c
        torch.cuda.empty_cache()

    # ---- same-box yardstick (not the product, outside every timed region): the es window through cuFFT -----------
    comparators = None
    if world == 1 and not args.no_cufft:
        try:
            dV = float(np.prod(dr))

            def cufft_es():
                return (torch.fft.irfftn(torch.fft.rfftn(potS) * torch.conj(torch.fft.rfftn(nrhoT)), s=potS.shape)[:, :, z_lo:z_hi] * dV)
            ref_c = cufft_es()
            mine = ops.density_overlap_window(potS, nrhoT, dr, z_lo, z_hi)
            err_c = float((mine - ref_c).abs().max() / ref_c.abs().max())
            del ref_c, mine
            ts = []
            for _ in range(3):
                a = ev(); cufft_es(); b = ev(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            comparators = {"cufft": {"ms_es": float(np.median(ts)), "value_es_only": N / (float(np.median(ts)) * 1e-3), "unit": UNIT,
                                     "rel_diff_vs_ours": err_c,
                                     "what": "torch.fft (cuFFT) float64: rfftn of both arrays, spectrum product with the conjugate, irfftn, "
                                             "window slice -- the straightforward library implementation of one interaction on the same GPU; "
                                             "compare with roofline.ms_es"}}
        except Exception as exc:                                   # noqa: BLE001 -- a comparator must never fail the bench
            comparators = {"cufft": {"error": repr(exc)[:200]}}
        torch.cuda.empty_cache()

    # ---- end to end: pinned host inputs in, results out, inside the timed region -----------
    e2e = None
    if not args.no_e2e:
        names = {0: ("rhoS", "rhoT", cfg.alpha), 1: ("potS", "nrhoT", 1.0)}
        dev_src = {"rhoS": rhoS, "potS": potS, "rhoT": rhoT, "nrhoT": nrhoT, "vdw": vdw}
        if slab_mode:
            # a rank uploads its x-slab of the four input grids and of the vdW window
            hosts = {k: torch.empty(v[sx0:sx1].shape, dtype=torch.float64).pin_memory().copy_(v[sx0:sx1]) for k, v in dev_src.items()}
            outs_h = [torch.empty((nx, ny, nzw), dtype=torch.float64).pin_memory() for _ in range(2)] if rank == 0 else []
        elif groups is not None:
            # distributed transform: a rank uploads only its x-slab of its interaction's two arrays
            need = set(names[my_which][:2])
            hosts = {k: torch.empty(dev_src[k][sx0:sx1].shape, dtype=torch.float64).pin_memory().copy_(dev_src[k][sx0:sx1])
                     for k in sorted(need)}
            hosts["vdw"] = torch.empty(vdw.shape, dtype=torch.float64).pin_memory().copy_(vdw)
            outs_h = [torch.empty((nx, ny, nzw), dtype=torch.float64).pin_memory() for _ in range(2)] if rank == 0 else []
        else:
            need = {"vdw"} | {n for (w, _, _) in jobs for n in names[w][:2]}
            hosts = {k: torch.empty(dev_src[k].shape, dtype=torch.float64).pin_memory().copy_(dev_src[k]) for k in sorted(need)}
            outs_h = [torch.empty(tuple(b.shape), dtype=torch.float64).pin_memory() for b in bufs]
        ni = my_ih - my_il
        out_rel = torch.empty((4, ni, ny, nz_relax), dtype=torch.float64).pin_memory()
        d_in = {k: torch.empty_like(v, device=dev) for k, v in hosts.items()}
        del rhoS, potS, rhoT, nrhoT, dev_src, inputs
        torch.cuda.empty_cache()

        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def e2e_slab_step(rec):
            e0 = ev()
            for k in hosts:
                d_in[k].copy_(hosts[k], non_blocking=True)
            src = {0: (d_in["rhoS"], d_in["rhoT"], cfg.alpha), 1: (d_in["potS"], d_in["nrhoT"], 1.0)}
            out = slab_step(None, src=src, vdw_slab=d_in["vdw"], sliced=True)     # static is rebuilt from this step's own windows
            for w, oh in zip((out["w_sr"], out["w_es"]) if rank == 0 else (), outs_h):
                oh.copy_(w, non_blocking=True)
            for q, key in enumerate(("E", "tip_x", "tip_y", "tip_z")):
                out_rel[q].copy_(out["r"][key], non_blocking=True)
            e_end = ev()
            torch.cuda.synchronize()
            if rec is not None:
                # sr+es: uploads -> both slabs finished; the rest: static, halo, relax, gathers, downloads
                rec.append((e0.elapsed_time(out["ev"][1]), out["ev"][1].elapsed_time(e_end)))

        def e2e_step(rec):
            if slab_mode:
                return e2e_slab_step(rec)
            e0 = ev()
            if groups is not None:
                na, nb, al = names[my_which]
                d_in[na].copy_(hosts[na], non_blocking=True)
                d_in[nb].copy_(hosts[nb], non_blocking=True)
                wins = parallel.sr_es_distributed(ops, d_in[na], d_in[nb], al, cfg.shape, dr, z_lo, z_hi, groups,
                                                  peer=not args.no_peer)
                for w, oh in zip(wins or [], outs_h):
                    oh.copy_(w, non_blocking=True)
            else:
                # uploads on a copy stream, results back on a third one: the sr compute and its download hide behind
                # the upload of the es inputs (PCIe is full duplex); an interaction starts when its two arrays have landed
                main = torch.cuda.current_stream()
                s_in.wait_stream(main)
                landed = []
                with torch.cuda.stream(s_in):
                    for (which, zl, zh) in jobs:
                        na, nb, al = names[which]
                        d_in[na].copy_(hosts[na], non_blocking=True)
                        d_in[nb].copy_(hosts[nb], non_blocking=True)
                        e_l = torch.cuda.Event(); e_l.record(s_in)
                        landed.append(e_l)
                for (which, zl, zh), buf, oh, e_l in zip(jobs, bufs, outs_h, landed):
                    na, nb, al = names[which]
                    main.wait_event(e_l)
                    if zh > zl:
                        ops.density_overlap_window(d_in[na], d_in[nb], dr, zl, zh, al, al, out=buf)
                    s_out.wait_stream(main)
                    with torch.cuda.stream(s_out):
                        oh.copy_(buf, non_blocking=True)
                main.wait_stream(s_out)
            e1 = ev()
            d_in["vdw"].copy_(hosts["vdw"], non_blocking=True)
            if world == 1:
                st = ops.build_static([(bufs[0], cfg.V), (bufs[1], 1.0), (d_in["vdw"], 1.0)])
            else:
                st = static
            rr = ops.relax_grid(st, cfg.kappa, cfg.rpivot, dr, nz_relax, tile=(my_il, my_ih, 0, ny))
            for q, key in enumerate(("E", "tip_x", "tip_y", "tip_z")):
                out_rel[q].copy_(rr[key], non_blocking=True)
            e2 = ev()
            torch.cuda.synchronize()
            if rec is not None:
                rec.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))

        e2e_step(None)
        rec = []
        barrier()
        for _ in range(args.steps):
            e2e_step(rec)
        barrier()
        tc, tr = max_over_ranks([float(np.mean([x[0] for x in rec])), float(np.mean([x[1] for x in rec]))])
        if slab_mode:
            h2d = 8 * (4 * N // G + Nw // G)
            d2h = 8 * (sum(int(o.numel()) for o in outs_h) + 4 * ni * ny * nz_relax)
        elif groups is not None:
            h2d = 8 * (2 * N // G + Nw)
            d2h = 8 * (sum(int(o.numel()) for o in outs_h) + 4 * ni * ny * nz_relax)
        else:
            h2d = 8 * (2 * len(jobs) * N + Nw)
            d2h = 8 * (sum(int(b.numel()) for b in bufs) + 4 * ni * ny * nz_relax)
        e2e = {"value": 2 * N / (tc * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_sr_es": tc, "relax_value": nx * ny * nz_relax / (tr * 1e-3), "relax_unit": "tip positions/s",
               "ms_relax": tr, "note": ("per rank: pinned host -> HBM copies of its x-slab of the four input grids and of the vdW window; "
                                        "sr, es, static (rebuilt from this step's own windows), halo exchange and relax as in the "
                                        "device-resident step; rank 0 copies both gathered windows back, every rank its relax tile; "
                                        "all inside the timed region; byte counts are rank 0's" if slab_mode else
                                        "per rank: pinned host -> HBM copies of its x-slab of its interaction's two input "
                                        "grids + vdw; rank 0 copies both windows back, every rank its relax tile; all "
                                        "inside the timed region; byte counts are rank 0's" if groups is not None else
                                        "per rank: pinned host -> HBM copies of the input grids of its interaction(s) "
                                        "+ vdw, its window slab(s) and relax tile copied back, all inside the timed "
                                        "region (uploads, compute and downloads on three streams so that they overlap); "
                                        "byte counts are per rank")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    dist_stages = None
    if dist_marks:
        acc = {}
        for rec_ in dist_marks:
            for (n0, e_a), (n1, e_b) in zip(rec_[:-1], rec_[1:]):
                acc.setdefault(n1, []).append(e_a.elapsed_time(e_b))
        dist_stages = {k: float(np.median(v)) for k, v in acc.items()}
        print("dist stage times per step (ms):", {k: [round(x, 3) for x in v] for k, v in acc.items()}, file=sys.stderr)
    peak, peak_src = measured_peak()
    balg = 8 * (2 * N + Nw)                       # per interaction
    t_corr = t_sr + t_es
    achieved = 2 * balg / (t_corr * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak / world,
            "traffic": (2 * profiled_traffic(args.workload) if world == 1 and profiled_traffic(args.workload) else None),
            "traffic_note": "dram bytes read+write of sr+es (2 x 5 launches), ncu --set full, profiles/traffic.json",
            "peak_source": peak_src,
            "kernel": ("dbspm_corr3d_dist_{fwd_peer,mid_peer,fin}_f64 = 5 launches per rank and interaction (axis_pass y_fwd with peer "
                       "stores | x_fwd, zfused, x_inv with peer stores | y_inv of the rank's x-slab); no gather inside the timed sr+es "
                       "region: every rank keeps its x-slab" if slab_mode else
                       "dbspm_corr3d_dist_{fwd_peer,mid,fin}_f64 = 5 launches per rank (axis_pass y_fwd with peer stores | "
                       "x_fwd, zfused, x_inv | y_inv of the rank's x-slab) + small all-to-all + gather" if groups is not None else
                       "dbspm_corr3d_f64 = 5 launches (axis_pass x_fwd, y_fwd, zfused, axis_pass y_inv, x_inv)"),
            "algorithmic_bytes_per_interaction": balg, "ms_sr_plus_es": t_corr,
            "ms_sr": t_sr if (world == 1 or slab_mode) else None, "ms_es": t_es if (world == 1 or slab_mode) else None,
            "note": "frac = algorithmic bytes of sr+es / (time * n_gpus * peak)"
                    + ("; es alone: %.4f of peak" % (balg / (t_es * 1e-3) / 1e9 / peak) if world == 1 else
                       "; sr then es, each over all ranks (x-slabs); time = inputs resident -> every rank holds its x-slab of the window"
                       if slab_mode else
                       "; sr and es run concurrently on disjoint rank groups, time includes the gather to rank 0")}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        c = cpu_reference_pass()
        cpu = {"value": c["pts_per_s"], "unit": UNIT, "cores": 1, "kind": "port", "sample": c["sample"],
               "relax_value": c["pos_per_s"], "relax_unit": "tip positions/s", "relax_cores": c["cores_relax"]}
    line = {
        "metric": METRIC, "value": 2 * N / (t_corr * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_all, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC.get(args.workload, args.workload), "grid": [nx, ny, nz],
                   "window_z": [z_lo, z_hi], "relax_shape": [nx, ny, nz_relax], "alpha": cfg.alpha, "V": cfg.V,
                   "kappa": cfg.kappa, "rpivot": cfg.rpivot, "tip": "noisy (non-compact, full-z path)",
                   "l2": "inputs are %.2f GB each, far larger than the 126 MB L2 (no flush needed)" % (8 * N / 1e9),
                   "parallelism": ((f"sr then es over all {G} ranks: x-slab decomposition of the transform (y pass whose stores go straight "
                                    f"into the owners' buffers over NVLink, x pass + z kernel + inverse x with the second exchange fused "
                                    f"the same way, inverse y); every rank keeps its x-slab of both windows, adds vdW, fetches a halo of "
                                    f"{halo} planes from its two neighbours and relaxes its x-tile in slab mode (bit-identical to the whole "
                                    f"window); windows and relax tiles gathered to rank 0 for the .npz, the window gather beside the relax kernel")
                                   if slab_mode else
                                   (f"sr | es on disjoint groups of {G} ranks; inside a group the transform is distributed "
                                    f"over x-slabs (y pass " + ("+ all-to-all through NCCL" if args.no_peer else
                                    "whose stores go straight into the owners' buffers over NVLink (symmetric memory)")
                                    + f", x pass + z kernel + inverse x, small all-to-all of the packed window rows back to "
                                    f"x-slabs, inverse y, gather of the real slabs into place on rank 0); x-tile x{world} (relax)")
                                   if groups is not None else
                                   (f"sr | es on disjoint rank groups, z-slab within a group ({world} ranks), "
                                    f"x-tile x{world} (relax), gather to rank 0") if world > 1 else "single GPU")},
        "relax": {"value": nx * ny * nz_relax / (relax_ms_max * 1e-3), "unit": "tip positions/s",
                  "ms": relax_ms_max, "mean_nfev": nfev_mean, "in_step_ms_static_relax_gather": t_rest, "in_step_stages_ms_max_over_ranks": rest_stages,
                  # the kernel is bound by L1 wavefronts, not HBM (compulsory HBM bytes are ~50 B/position):
                  # algorithmic stencil traffic = scipy's nfev x 64 points x 8 B per position, against the
                  # nominal L1 load bandwidth (SMs x 128 B/clk x max SM clock)
                  "stencil_gbs": nx * ny * nz_relax * nfev_mean * 512 / (relax_ms_max * 1e-3) / 1e9,
                  "l1_nominal_gbs": world * 148 * 128 * 1.965,
                  "hbm_bytes_per_position": 8.0 * nzw / nz_relax + 56.0},
        "stages_ms": stages, "dist_stages_ms": dist_stages, "dist_check_rel_err": dist_check,
        "dist_relax_tile_bit_identical": relax_check, "compact_tip": compact, "roofline": roof, "comparators": comparators,
        "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": n_launch, "clocks": clk.summary(), "wall_s_timed_region": t_wall,
    }
    # numbers first, prose last (a truncated tail of the line keeps the free-text config, a truncated head the figures)
    order = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "dist_check_rel_err", "dist_relax_tile_bit_identical", "dist_stages_ms", "relax", "roofline", "e2e",
             "gpu_launches", "clocks", "stages_ms", "comparators", "cpu_baseline", "compact_tip", "wall_s_timed_region", "config"]
    line = {k: line[k] for k in order if k in line} | {k: v for k, v in line.items() if k not in order}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large_slab", choices=sorted(WORKLOAD_DESC))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="distributed transform: NCCL all-to-all instead of the fused peer-store exchange")
    ap.add_argument("--no-dist", action="store_true", help="N >= 4: replicated inputs + z-slab split instead of the distributed transform")
    ap.add_argument("--legacy-groups", action="store_true",
                    help="N >= 4: round-1 plan (an sr group and an es group, windows gathered to rank 0 and broadcast back for relax)")
    ap.add_argument("--no-compact", action="store_true", help="skip the compact-tip (z-pruned path) measurement")
    ap.add_argument("--no-reference-full", dest="reference_full", action="store_false",
                    help="--impl reference: skip the one full-size sr+es pass (about 3 minutes of numpy fftn), bounded sample only")
    ap.add_argument("--no-cufft", action="store_true", help="skip the cuFFT (torch.fft) comparator")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "sweep16":
            args.workload = "pyridine"
        return run_reference_arm(args)
    if args.workload == "sweep16":
        return run_sweep_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    raise SystemExit(main())
