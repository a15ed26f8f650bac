"""Force -> frequency-shift post-processing on the GPU (reference: pydbspm/tools.py:5-32 and
SPMGrid.get_gradient, pydbspm/SPMGrid.py:375-379).  These are the "Delta-f images" whose
tolerance north_star states; they are streaming epilogues of relax."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .interactions import _device, _stream_ptr, _to_dev, force_gradient


def giessibl_weights(dz=0.075, n=10):
    """Weights and prefactor of get_fs (tools.py:24-29): derivative of the semicircle sqrt(1-x^2)
    sampled on n+1 points, with the small-n correction."""
    x = np.linspace(-1, 1, int(n + 1))
    y = np.sqrt(1 - x * x)
    dy = (y[1:] - y[:-1]) / (dz * n)
    fpi = (n - 2) ** 2
    prefactor = (1 + fpi * (2 / np.pi)) / (fpi + 1)
    return dy, prefactor


def get_fs(F, dz=0.075, k0=1800, f0=30300, n=10, conv=16.0217656):
    """Drop-in for pydbspm.tools.get_fs: vertical force (nx,ny,nz) -> frequency shift
    (nx,ny,nz-n+1) = -prefactor * convolve_valid(F, dy) * conv * f0 / k0 along z."""
    on_device = isinstance(F, torch.Tensor) and F.is_cuda
    dev = F.device if on_device else _device()
    t = _to_dev(F, dev)
    dy, prefactor = giessibl_weights(dz, n)
    nx, ny, nz = (int(s) for s in t.shape)
    w = torch.from_numpy(np.ascontiguousarray(dy)).to(dev)
    out = torch.empty((nx, ny, nz - len(dy) + 1), dtype=torch.float64, device=dev)
    scale = -prefactor * conv * f0 / k0
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbspm_convz_valid_f64(t.data_ptr(), nx, ny, nz, w.data_ptr(), len(dy), float(scale),
                                                    out.data_ptr(), _stream_ptr(dev)))
    return out if on_device else out.cpu().numpy()


def get_df_image(E, dz, **kw):
    """E (relax energy grid) -> Delta-f: get_fs(np.gradient(E, -dz, axis=2), dz, ...)."""
    return get_fs(force_gradient(E, dz), dz=dz, **kw)
