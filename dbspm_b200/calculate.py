"""DensityCalculator: owner of the sample / tip arrays and of the scan window.

Same attributes and methods as the reference's pydbspm/calculate.py:5-72 (load_sample,
load_tip, set_sample, set_tip, set_calculation_grid; rhoS, potS, gridS, rhoT, nrhoT, gridT,
grid, nidx).  Added: the arrays are staged once in HBM (`device_array`) so that sr and es do
not pay the host->device copy twice, and `z_window` exposes the window as plain ints for
the C ABI."""
from __future__ import annotations

import numpy as np
import torch

from .grid import Grid, get_grid_from_params
from .npzio import read_npz


class DensityCalculator:
    def __init__(self, params=None, device=None):
        self.params = params
        self.sample = self.tip = None
        self.rhoS = self.potS = self.gridS = None
        self.rhoT = self.nrhoT = self.gridT = None
        self.grid = self.nidx = None
        self.stm = self.stms = self.stmp = self.tippos = None
        self.device = device
        self._dev_cache = {}
        self._pinned = {}

    # -- reference API ---------------------------------------------------------------------
    def set_sample(self, gridS):
        self.gridS, self.rhoS, self.potS = gridS, gridS.rhoS, gridS.potS
        self._dev_cache.clear()

    def set_tip(self, gridT):
        self.gridT, self.rhoT, self.nrhoT = gridT, gridT.rhoT, gridT.nrhoT
        self._dev_cache.clear()

    def _npz(self, given, attr, fallback):
        if given:
            return given
        if self.params:
            return getattr(self.params, attr) / fallback.split("/")[-1]
        return fallback

    def _load(self, path):
        """`.npz` -> mapping of numpy arrays; the big ones are views of pinned host memory (kept alive in
        self._pinned) so that device_array() uploads them at full PCIe speed without a staging copy."""
        data = read_npz(path)
        self._pinned.update({(str(path), k): v for k, v in data.items() if isinstance(v, torch.Tensor)})
        return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in data.items()}

    def load_sample(self, sample_dir=None):
        self.sample = self._load(self._npz(sample_dir, "SAMPLE", "sample/sample.npz"))
        self.rhoS, self.potS = self.sample["rhoS"], self.sample["potS"]
        self.gridS = Grid(self.rhoS.shape, cell=self.sample["cell"], positions=self.sample["positions"],
                          numbers=self.sample["numbers"])
        self._dev_cache.clear()

    def load_tip(self, tip_dir=None):
        self.tip = self._load(self._npz(tip_dir, "TIP", "tip/tip.npz"))
        self.rhoT, self.nrhoT = self.tip["rhoT"], self.tip["nrhoT"]
        self.gridT = Grid(self.rhoT.shape, cell=self.tip["cell"], positions=self.tip["positions"],
                          numbers=self.tip["numbers"])
        self._dev_cache.clear()

    def set_calculation_grid(self):
        self.grid, self.nidx = get_grid_from_params(
            self.params, self.gridS, nidx=True, rpivot=self.params.RPIVOT, zpad=self.params.ZPAD,
            numbers=self.gridS.numbers, positions=self.gridS.positions)

    # -- device staging --------------------------------------------------------------------
    def device_array(self, name: str) -> torch.Tensor:
        """rhoS / potS / rhoT / nrhoT as a float64 CUDA tensor, uploaded once."""
        if name not in self._dev_cache:
            host = getattr(self, name)
            if host is None:
                raise ValueError(f"{name} is not loaded")
            if isinstance(host, torch.Tensor):
                t = host.to(dtype=torch.float64)
            else:
                t = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64))
            dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            self._dev_cache[name] = t.to(dev).contiguous()
        return self._dev_cache[name]

    @property
    def z_window(self):
        """(z_lo, z_hi) of the scan window in sample-grid plane indices."""
        if self.nidx is None:
            raise ValueError("call set_calculation_grid() first")
        sl = self.nidx[2]
        return int(sl.start), int(sl.stop)
