"""`dbspm <steps...> -i input.in` for the sr / es / relax steps on B200.

Same step names and aliases (setup_calculation.py), same flags and precedence
default < input file < command line (dbspm:216-231), same input files (SAMPLE/sample.npz,
TIP/tip.npz, LABEL_*.npz) and the same output file names and `.npz` keys as the reference
driver (dbspm:319-355, :468-620, :644-667).  `brstm` (the tricubic gather at the relaxed tip
positions) runs too; steps that are not on the GPU path (sample tip grid vdw stm) are refused with a message instead of being silently skipped: run those with
the reference and point this program at their outputs.

Multi-GPU: launch under torchrun (one process per GPU); sr/es shard by window z-slab and
relax by lateral x-tile (parallel.py), rank 0 writes the files."""
from __future__ import annotations

from argparse import ArgumentParser, Namespace
from datetime import datetime as dt

import numpy as np
import torch

from . import interactions as ops
from . import parallel
from .calculate import DensityCalculator
from .grid import get_grid_from_params
from .input_parser import default_params, parse_params
from .setup_calculation import gpu_steps, setup_calc


def _parser() -> ArgumentParser:
    p = ArgumentParser(prog="DBSPM", description="Density-based SPM simulation: sr / es / relax steps on B200. "
                       "Command-line arguments override parameters in the input file.")
    p.add_argument("steps", nargs="*", default=["fdbm"], help="Steps to calculate.")
    p.add_argument("--components", dest="COMPONENTS", nargs="*",
                   help="Labels of the interactions to include in the static potential.")
    p.add_argument("-i", "--input", dest="INPUT", type=str, help="Input file with simulation parameters.")
    p.add_argument("-l", "--label", dest="LABEL", type=str, help="Label for the output files.")
    p.add_argument("-s", "--sample", dest="SAMPLE", type=str, help="Sample directory (holds sample.npz).")
    p.add_argument("-t", "--tip", dest="TIP", type=str, help="Tip directory (holds tip.npz).")
    p.add_argument("-a", "--alpha", dest="ALPHA", type=float, help="Exponent for Pauli-repulsion (SR) overlap.")
    p.add_argument("-V", dest="V", type=float, help="Pauli-repulsion (SR) scaling parameter.")
    p.add_argument("-k", "--kappa", dest="K", type=float, help="Probe torsion spring constant. [eV/rad]")
    p.add_argument("-r", "--rpivot", dest="RPIVOT", type=float, help="Length of the probe lever. [Ang]")
    p.add_argument("--min-method", dest="MINMETHOD", type=str,
                   help="Minimization method.  Only 'Powell' (the reference's default, restated step for step on the GPU) is "
                        "available; any other scipy method name raises NotImplementedError.")
    p.add_argument("-b", "--bias", action="store", type=float, default=0.5, help="STM bias (V) (brstm step).")
    p.add_argument("-w", "--weight", action="store", type=float, default=0.5,
                   help="weight of the s orbital in the BRSTM calculation.")
    return p


def _merge_params(options: Namespace, say) -> Namespace:
    from pathlib import Path
    if options.INPUT:
        say(f"\n> Reading parameters from {options.INPUT}.")
        file_params = parse_params(options.INPUT).__dict__
    else:
        say("\n> No input file provided, using default parameters.")
        file_params = {}
    merged = dict(default_params)
    merged.update(file_params)
    for k, v in vars(options).items():
        if k not in ("steps", "bias", "weight") and v is not None:
            merged[k] = Path(v) if k in ("SAMPLE", "TIP") else v
    for k in default_params:
        src = "c" if getattr(options, k, None) is not None else ("i" if file_params.get(k) is not None else "d")
        say(f"{k:10s} = {str(merged[k]):20s} ({src})")
    # LABEL stays the raw capture of the parser (the reference's regex keeps blanks before a `#` comment, so
    # `LABEL = foo  # c` names its files 'foo  _es.npz'): files written here carry the reference's names
    merged.setdefault("LABEL", "dbspm")
    return Namespace(**merged)


def _existing(name: str) -> str:
    """A file the reference (or this program) wrote under LABEL; tolerates a LABEL with stray blanks on either side."""
    import os
    if os.path.exists(name):
        return name
    alt = name.replace(" ", "")
    return alt if os.path.exists(alt) else name


def main(argv=None) -> int:
    t_start = dt.now()
    rank, world, device = parallel.init_from_env()
    say = (lambda *a, **k: print(*a, **{**k, "flush": True})) if rank == 0 else (lambda *a, **k: None)
    options = _parser().parse_args(argv)
    say("Density-based SPM (B200 hot path: sr / es / relax)")
    say(f"Started on {t_start:%Y-%m-%d %H:%M:%S}")
    say(f"Calculating on {world} GPU process(es).")
    params = _merge_params(options, say)
    steps = setup_calc(options.steps)
    refused = [k for k, v in vars(steps).items() if v and k not in gpu_steps]
    if refused:
        say("\nThese steps are outside the GPU hot path and must be run with the reference "
            f"implementation: {', '.join(refused)}")
        return 2
    steptime = {}
    calc = DensityCalculator(params, device=device)

    def ensure_inputs(sample=True, tip=True):
        if sample and calc.gridS is None:
            calc.load_sample()
        if tip and calc.gridT is None:
            calc.load_tip()
        if calc.grid is None:
            calc.set_calculation_grid()

    def run_overlap(label, a_name, b_name, alpha, filename):
        ensure_inputs()
        z_lo, z_hi = calc.z_window
        nx, ny, _ = calc.rhoS.shape
        a, b = calc.device_array(a_name), calc.device_array(b_name)
        # a tip array that is exactly zero outside a z-range (a cut-off or synthetic tip) takes the pruned
        # transform; a CHGCAR tip has a noise floor everywhere and the support comes back as the whole cell
        support = ops.tip_zsupport(b) if b.shape[2] % 2 == 0 else None
        if support is not None and support[1] >= b.shape[2]:
            support = None

        def slab(zl, zh):
            return ops.density_overlap_window(a, b, calc.gridS.dr, zl, zh, alpha, alpha, tip_support=support)

        nz_problem = a.shape[2]
        if world > 1:
            cut = ops.pruned_cut(a.shape[2], z_lo, z_hi, support)
            nz_problem = cut[1] if cut is not None else a.shape[2]
        if world > 1 and ops.dist_supported((a.shape[0], a.shape[1], nz_problem), world):
            # the transform itself is distributed: this rank transforms its x-slab (only the z-ranges the window can see
            # when the tip is compact), exchange, gather to rank 0
            x0, x1 = parallel.x_slab(nx, world, rank)
            wins = parallel.sr_es_distributed(ops, a[x0:x1], b[x0:x1], alpha, tuple(a.shape), calc.gridS.dr, z_lo, z_hi,
                                              [list(range(world))], tip_support=support)
            win = wins[0] if rank == 0 else None
        else:
            win = parallel.corr3d_sharded(slab, nx, ny, z_lo, z_hi, device)
        if rank == 0:
            calc.grid.set_density(label, win.cpu().numpy(), safe=False)
            say(f"- Saving {filename}")
            calc.grid.save_density(label, filename=filename)
        # the reference calls comm.Barrier() after every step (dbspm:335,354): the next step of the same invocation
        # (relax) reads this file on every rank, so nobody may run ahead of rank 0's write
        parallel.barrier()

    if steps.sr:
        say("> Calculating SR interaction.")
        t0 = dt.now()
        run_overlap("sr", "rhoS", "rhoT", params.ALPHA, params.LABEL + f"_sr_a{params.ALPHA:.2f}.npz")
        steptime["sr"] = dt.now() - t0
        say("< done.\n")
    if steps.es:
        say("> Calculating ES interaction.")
        t0 = dt.now()
        run_overlap("es", "potS", "nrhoT", 1.0, params.LABEL + "_es.npz")
        steptime["es"] = dt.now() - t0
        say("< done.\n")
    if steps.relax:
        say("Relaxing probe on static potential.")
        t0 = dt.now()
        ensure_inputs(tip=False)
        comps = []
        for name in params.COMPONENTS:
            name = str(name).lower()
            if name == "sr":
                comps.append((np.load(_existing(params.LABEL + f"_sr_a{params.ALPHA:.2f}.npz"))["sr"], params.V))
            else:
                comps.append((np.load(_existing(params.LABEL + f"_{name}.npz"))[name], 1.0))
        static = ops.build_static(comps, device=device)
        gridS = calc.gridS
        relax_grid, _, nbox = get_grid_from_params(params, gridS, nidx=True, nbox=True, numbers=gridS.numbers,
                                                   positions=gridS.positions)
        nz_relax = int(nbox[1, 2] - nbox[0, 2])
        dR = None if gridS.ortho else relax_grid.dR
        field = ops.TricubicField(static)
        say("> Energy minimization." if gridS.ortho else "> Energy minimization for non orthogonal cell.")

        def tile(i_lo, i_hi):
            r = ops.relax_grid(field, params.K, params.RPIVOT, gridS.dr, nz_relax, dR=dR,
                               tile=(i_lo, i_hi, 0, field.shape[1]), method=params.MINMETHOD)
            return {k: r[k] for k in ("E", "tip_x", "tip_y", "tip_z")}

        res = parallel.relax_sharded(tile, field.shape[0], device)
        say("< done.\n")
        if rank == 0:
            E = res["E"].cpu().numpy()
            say("F_max = {:.4f} pN".format(1602 * np.gradient(E, -relax_grid.dr[2], axis=2).max()))
            fname = f"relax_{params.LABEL}_k{params.K:.4f}_a{params.ALPHA:.2f}_V{params.V:.2f}.npz"
            say(f"> Saving {fname} array...")
            for key in ("E", "tip_x", "tip_y", "tip_z"):
                relax_grid.set_density(key, res[key].cpu().numpy(), safe=False)
            relax_grid.save_density("E", "tip_x", "tip_y", "tip_z", filename=fname)
        parallel.barrier()                                   # dbspm:619
        steptime["relax"] = dt.now() - t0
    if steps.brstm:
        # dbspm:644-667: s/p-wave STM signal gathered at the relaxed tip-apex positions.  The stm file
        # (stm_LABEL_V*.npz, written by the reference's `stm` step) and the relax file are read; the
        # tricubic gather runs on the GPU.  zref comes from the stm file (the reference re-derives it
        # from the CONTCAR through ASE).
        if rank == 0:
            from scipy.ndimage import gaussian_filter
            from .grid import interpolate_data, read_grid
            say("Calculating BRSTM image.")
            t0 = dt.now()
            rel_name = f"relax_{params.LABEL}_k{params.K:.4f}_a{params.ALPHA:.2f}_V{params.V:.2f}.npz"
            say("Getting tip position from " + rel_name)
            relax = read_grid(rel_name)
            stm = read_grid(f"stm_{params.LABEL}_V{options.bias:.3f}.npz")
            ws_, wp_ = options.weight, 1 - options.weight
            stm.set_density("sp", ws_ * stm.s + gaussian_filter(wp_ * stm.p, sigma=2, mode="wrap"))
            out = interpolate_data(relax, stm, "sp", params.RPIVOT)
            out.save_density("data", filename=f"brstm_{params.LABEL}_V{options.bias:.3f}_ws{ws_:.2f}.npz")
            say("BRSTM calculated.")
            steptime["brstm"] = dt.now() - t0
        parallel.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t_end = dt.now()
    say(f"Finished on {t_end:%Y-%m-%d %H:%M:%S}")
    say(f"Time elapsed: {t_end - t_start}")
    say(f"{'Steps':7s} | time (h:m:s)")
    say("-" * 24)
    for k, v in steptime.items():
        say(f"{k:7s} | {v}")
    say("Done.")
    parallel.shutdown()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
