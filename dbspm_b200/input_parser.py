"""`input.in` parser (host side).  Same file format, keys, types, defaults and quirks as the
reference's pydbspm/input_parser.py:7-119, restated as one typed table:

  * one `KEY = value` per line, value runs to `#`, `|`, `^` or end of line (regex of :80);
  * typed keys are converted (path / str / bool / float / float-array); a boolean is True only
    for the exact text "True"; untyped keys (e.g. LABEL) stay raw strings;
  * BBOX0 + BBOXF, when both present, are stacked into BBOX;
  * `&SITES` / `&VASP` blocks and `$KEY` lines are ignored here (they feed the DFT steps).
"""
from __future__ import annotations

import re
from argparse import Namespace
from pathlib import Path

import numpy as np

# key -> (kind, default)
_SPEC = {
    "TIP": ("path", Path("./tip")), "SAMPLE": ("path", Path("./sample")),
    "GRID": ("path", Path("grid.npz")), "DFT": ("path", Path("./DFT/Fz")),
    "PSEUDOS": ("str", "potpaw_PBE"), "CMD": ("str", "mpirun vasp_gam"), "CALC": ("str", "vasp"),
    "SAMPLERHO": ("str", None), "SAMPLEPOT": ("str", None), "TIPRHO": ("str", None),
    "TIPNRHO": ("str", None), "MINMETHOD": ("str", "Powell"),
    "DFTCHG": ("bool", True), "DFTPOT": ("bool", True), "TIPGRID": ("bool", True),
    "SAMPLEGRID": ("bool", True), "POTFILTER": ("bool", False), "CENTER": ("bool", False),
    "FIT": ("bool", False), "FITV0": ("bool", False), "REFIT": ("bool", False), "REAL": ("bool", True),
    "MPI": ("bool", False), "RELAX": ("bool", True), "OPT": ("bool", False),
    "ALPHA": ("float", 1.08), "V": ("float", 42.91), "Z0": ("float", 2.8), "ZF": ("float", 4.0),
    "ZREF": ("float", None), "ZPAD": ("float", 1.0), "RPIVOT": ("float", 3.02), "K": ("float", 0.2),
    "ZVAL": ("array", None), "BBOX0": ("array", None), "BBOXF": ("array", None), "BBOX": ("array", None),
    "ARANGE": ("array", [1.0, 1.12, 0.01]), "XLIM": ("array", [0.0, 1.0]), "YLIM": ("array", [0.0, 1.0]),
    "COMPONENTS": ("array", ["sr", "es", "vdw"]),
    "COMM": ("internal", None),
}

default_params = {k: d for k, (_, d) in _SPEC.items()}

_LINE = re.compile(r"^(\w+)\s*=\s*([^#|^\n]+)", re.M)

_CONVERT = {
    "path": Path,
    "str": str,
    "bool": lambda s: s == "True",
    "float": float,
    "array": lambda s: np.array(s.split(), dtype=float),
}


def parse_params(path: str) -> Namespace:
    with open(path) as fh:
        found = dict(_LINE.findall(fh.read()))
    for key, (kind, default) in _SPEC.items():
        if kind == "internal":
            continue
        found[key] = _CONVERT[kind](found[key]) if key in found else default
    if isinstance(found["BBOX0"], np.ndarray) and isinstance(found["BBOXF"], np.ndarray):
        found["BBOX"] = np.stack([found["BBOX0"], found["BBOXF"]])
    return Namespace(**found)
