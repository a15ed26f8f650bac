"""Step names and aliases of the `dbspm` command line (reference: pydbspm/setup_calculation.py:4-69).

The step set and every alias are kept; only sr / es / relax are executed by this package,
the other steps are recognised so that mixed command lines are diagnosed instead of ignored."""
from __future__ import annotations

from argparse import Namespace

STEP_NAMES = ("sample", "tip", "grid", "sr", "es", "vdw", "relax", "stm", "brstm")
_STATIC = ("sr", "es", "vdw")
_ALIASES = {
    "prepdft": ("sample", "tip"),
    "prep": ("grid",),
    "fullprep": ("sample", "tip", "grid"),
    "pauli": ("sr",),
    "dens": ("sr", "es"), "density_interactions": ("sr", "es"),
    "disp": ("vdw",), "d3": ("vdw",),
    "static": _STATIC, "all": _STATIC,
    "afm": _STATIC + ("relax",), "fdbm": _STATIC + ("relax",),
    "fromdft": ("grid",) + _STATIC + ("relax",),
    "fullauto": ("sample", "tip", "grid") + _STATIC + ("relax",),
    "dft+afm": ("sample", "tip", "grid") + _STATIC + ("relax",),
}

mpirequired = ["vdw", "relax"]
#: steps this package runs on the GPU; the rest stay with the reference implementation
gpu_steps = ("sr", "es", "relax", "brstm")


def setup_calc(words) -> Namespace:
    flags = dict.fromkeys(STEP_NAMES, False)
    for w in words:
        if w in flags:
            flags[w] = True
        elif w in _ALIASES:
            for s in _ALIASES[w]:
                flags[s] = True
        else:
            print("Unknown option: {}\nIgnoring...".format(w))
    return Namespace(**flags)
