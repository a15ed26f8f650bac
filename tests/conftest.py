import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def overlap_golden():
    return np.load(os.path.join(GOLDEN, "overlap_reference.npz"))


@pytest.fixture(scope="session")
def relax_golden():
    return np.load(os.path.join(GOLDEN, "relax_reference.npz"))


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Emu:
    """ctypes face of tests/emu/libdbspm_emu.so: the CUDA kernel bodies compiled for the CPU."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "emu")
        subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(d, "libdbspm_emu.so"))
        L.emu_corr3d_f64.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, _dp, C.c_int]
        L.emu_corr3d_pruned_f64.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
        L.emu_corr3d_dist_f64.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, _dp, C.c_int, C.c_int]
        L.emu_corr3d_dist_pruned_f64.argtypes = [_dp, _dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int]
        L.emu_dist_row_of.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong]
        L.emu_pow_fast.argtypes = [_dp, _dp, C.c_longlong, C.c_double]
        L.emu_fft_lines.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int]
        L.emu_fft_lines_t.argtypes = [_dp, _dp, C.c_int, C.c_int, C.c_int]
        L.emu_relax_powell_f64.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_double, C.c_double,
                                           C.c_int, _dp, _dp, _ip, C.c_int]
        L.emu_tricubic.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.emu_tricubic.restype = C.c_double
        self.L = L

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(_dp)

    def corr3d(self, a, b, dr, z_lo=0, z_hi=None, alpha0=1.0, alpha1=1.0, tshift=-1):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        nx, ny, nz = a.shape
        z_hi = nz if z_hi is None else z_hi
        out = np.empty((nx, ny, z_hi - z_lo))
        rc = self.L.emu_corr3d_f64(self._p(a), self._p(b), alpha0, alpha1, float(np.prod(dr)), nx, ny, nz,
                                   z_lo, z_hi, self._p(out), tshift)
        if rc:
            raise RuntimeError(f"emu_corr3d_f64 -> {rc}")
        return out

    def corr3d_dist(self, a, b, dr, z_lo, z_hi, G, alpha0=1.0, alpha1=1.0, tshift=-1):
        """the x-slab distributed correlation with G ranks run one after the other"""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        nx, ny, nz = a.shape
        out = np.empty((nx, ny, z_hi - z_lo))
        rc = self.L.emu_corr3d_dist_f64(self._p(a), self._p(b), alpha0, alpha1, float(np.prod(dr)), nx, ny, nz,
                                        z_lo, z_hi, self._p(out), G, tshift)
        if rc:
            raise RuntimeError(f"emu_corr3d_dist_f64 -> {rc}")
        return out

    def corr3d_dist_pruned(self, a, b, dr, z_lo, z_hi, s_lo, s_len, G, alpha0=1.0, alpha1=1.0, tshift=-1):
        """-> (window, z length actually transformed) of the distributed step with a compact tip"""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        nx, ny, nz = a.shape
        out = np.empty((nx, ny, z_hi - z_lo))
        rc = self.L.emu_corr3d_dist_pruned_f64(self._p(a), self._p(b), alpha0, alpha1, float(np.prod(dr)), nx, ny, nz,
                                               z_lo, z_hi, s_lo, s_len, self._p(out), G, tshift)
        if rc < 0:
            raise RuntimeError(f"emu_corr3d_dist_pruned_f64 -> {rc}")
        return out, rc

    def pow_fast(self, x, alpha):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self.L.emu_pow_fast(self._p(x), self._p(out), x.size, float(alpha))
        return out

    def corr3d_pruned(self, a, b, dr, z_lo, z_hi, s_lo, s_len, alpha0=1.0, alpha1=1.0, tshift=-1):
        """-> (window, z length actually transformed)"""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        nx, ny, nz = a.shape
        out = np.empty((nx, ny, z_hi - z_lo))
        rc = self.L.emu_corr3d_pruned_f64(self._p(a), self._p(b), alpha0, alpha1, float(np.prod(dr)), nx, ny, nz,
                                          z_lo, z_hi, s_lo, s_len, self._p(out), tshift)
        if rc < 0:
            raise RuntimeError(f"emu_corr3d_pruned_f64 -> {rc}")
        return out, rc

    def fft_lines(self, x, direction, tshift, transposed=False):
        n, T = x.shape
        xin = np.ascontiguousarray(x.astype(np.complex128)).view(np.float64)
        out = np.empty_like(xin)
        fn = self.L.emu_fft_lines_t if transposed else self.L.emu_fft_lines
        rc = fn(self._p(xin), self._p(out), n, tshift, direction)
        if rc < 0:
            raise RuntimeError(f"emu_fft_lines -> {rc}")
        return out.view(np.complex128).reshape(n, T)

    def relax(self, s, tile, nz, kappa, rpivot, dr, dR=None, maxfev=2000, xtol=1e-4, ftol=1e-4, transposed=0):
        s = np.ascontiguousarray(s, dtype=np.float64)
        nx, ny, nzc = s.shape
        i_lo, i_hi, j_lo, j_hi = tile
        ni, nj = i_hi - i_lo, j_hi - j_lo
        etp = np.empty((ni, nj, nz, 3))
        cart = np.empty((3, ni, nj, nz))
        nfev = np.zeros((ni, nj, nz), dtype=np.int32)
        drv = np.ascontiguousarray(dr, dtype=np.float64)
        dRv = None if dR is None else np.ascontiguousarray(dR, dtype=np.float64).reshape(9)
        rc = self.L.emu_relax_powell_f64(self._p(s), nx, ny, nzc, i_lo, i_hi, j_lo, j_hi, nz, kappa, rpivot,
                                         self._p(drv), None if dRv is None else self._p(dRv), xtol, ftol, maxfev,
                                         self._p(etp), self._p(cart), nfev.ctypes.data_as(_ip), transposed)
        if rc:
            raise RuntimeError(f"emu_relax_powell_f64 -> {rc}")
        return etp, cart, nfev

    def relax_slab(self, slab, nx, x_base, tile, nz, kappa, rpivot, dr, dR=None, maxfev=2000, transposed=0):
        """x-slab mode: `slab` holds the planes x_base ... (mod nx) of the window; `tile` in window coordinates."""
        s = np.ascontiguousarray(slab, dtype=np.float64)
        nx_local, ny, nzc = s.shape
        i_lo, i_hi, j_lo, j_hi = tile
        ni, nj = i_hi - i_lo, j_hi - j_lo
        etp = np.empty((ni, nj, nz, 3)); cart = np.empty((3, ni, nj, nz)); nfev = np.zeros((ni, nj, nz), dtype=np.int32)
        drv = np.ascontiguousarray(dr, dtype=np.float64)
        dRv = None if dR is None else np.ascontiguousarray(dR, dtype=np.float64).reshape(9)
        f = self.L.emu_relax_powell_slab_f64
        f.argtypes = [_dp] + [C.c_int] * 10 + [C.c_double, C.c_double, _dp, _dp, C.c_double, C.c_double, C.c_int, _dp, _dp, _ip, C.c_int]
        rc = f(self._p(s), nx, ny, nzc, x_base, nx_local, i_lo, i_hi, j_lo, j_hi, nz, kappa, rpivot, self._p(drv),
               None if dRv is None else self._p(dRv), 1e-4, 1e-4, maxfev, self._p(etp), self._p(cart), nfev.ctypes.data_as(_ip), transposed)
        if rc:
            raise RuntimeError(f"emu_relax_powell_slab_f64 -> {rc}")
        return etp, cart, nfev

    def corr3d_direct(self, a, pts, w, dr, z_lo, z_hi, alpha0=1.0):
        """direct real-space correlation with a tip given as support points `pts` (K,3 indices) and weights w (= B^alphaB)"""
        a = np.ascontiguousarray(a, dtype=np.float64)
        nx, ny, nz = a.shape
        pts = np.ascontiguousarray(pts, dtype=np.int32)
        sx, sy, sz = (np.ascontiguousarray(pts[:, q]) for q in range(3))
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.empty((nx, ny, z_hi - z_lo))
        f = self.L.emu_corr3d_direct_f64
        f.restype = C.c_longlong
        f.argtypes = [_dp, C.c_double, C.c_double] + [C.c_int] * 6 + [_ip, _ip, _ip, _dp, _dp]
        rc = f(self._p(a), alpha0, float(np.prod(dr)), nx, ny, nz, z_lo, z_hi, len(w), sx.ctypes.data_as(_ip), sy.ctypes.data_as(_ip),
               sz.ctypes.data_as(_ip), self._p(w), self._p(out))
        if rc < 0:
            raise RuntimeError(f"emu_corr3d_direct_f64 -> {rc}")
        return out, int(rc)

    def relax_patch_columns(self, ni, nj, foot_x):
        f = self.L.emu_relax_patch_columns
        f.restype = C.c_longlong
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_longlong]
        cap = (ni + 32) * (nj + 32) * 2 + 64
        cols = np.empty(cap, dtype=np.int64)
        n = f(ni, nj, foot_x, cols.ctypes.data_as(C.POINTER(C.c_longlong)), cap)
        assert n <= cap
        return cols[:n]

    def tricubic(self, g, x, y, z):
        g = np.ascontiguousarray(g, dtype=np.float64)
        return self.L.emu_tricubic(self._p(g), *g.shape, x, y, z)


@pytest.fixture(scope="session")
def emu():
    return Emu()
