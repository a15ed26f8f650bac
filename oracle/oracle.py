"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under dbspm_b200/ does; the
product path raises if its CUDA library is missing instead of falling back here.

Contents (reference = /root/reference, read-only, absent on the GPU box):
  density_overlap_fft   numpy restatement of pydbspm/interactions.py:9-22
  corr_direct           the same quantity as an explicit periodic sum (C, small grids)
  Tricubic              restatement of the third-party `tricubic` package's interface as the
                        reference uses it (dbspm:562, interactions.py:59): Tricubic(list|array,
                        shape).ip([x, y, z]); Lekien-Marsden, periodic, index units
  relax_tip_z_scipy     restatement of interactions.py:70-86 on top of the *installed* scipy
                        (what the reference actually runs, Python overheads included)
  relax_tile            C restatement of the same (scipy's Powell restated in C); twin=True swaps in the CUDA
                        kernel's own stencil arithmetic and sin/cos (bit-exact target for the GPU tests)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dbl_p = C.POINTER(C.c_double)
_int_p = C.POINTER(C.c_int)
_long_p = C.POINTER(C.c_long)


def build(force: bool = False) -> str:
    """Compile liboracle.so next to this file (gcc only)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dbspm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_corr_direct.argtypes = [_dbl_p, _dbl_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.c_int, C.c_int, _dbl_p]
        L.oracle_corr_direct.restype = None
        L.tc_create.argtypes = [_dbl_p, C.c_int, C.c_int, C.c_int]
        L.tc_create.restype = C.c_void_p
        L.tc_destroy.argtypes = [C.c_void_p]
        L.tc_destroy.restype = None
        L.tc_ip.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.tc_ip.restype = C.c_double
        L.tc_ip_catmull.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.tc_ip_catmull.restype = C.c_double
        L.oracle_relax_column.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_double, C.c_double,
                                          _long_p, C.c_int, _dbl_p, _dbl_p, C.c_double, C.c_double,
                                          C.c_int, _dbl_p, _int_p]
        L.oracle_relax_column.restype = None
        L.oracle_relax_tile.argtypes = [_dbl_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_double, C.c_double, _dbl_p, _dbl_p,
                                        C.c_double, C.c_double, C.c_int, _dbl_p, _int_p]
        L.oracle_relax_tile.restype = None
        L.oracle_relax_tile_twin.argtypes = L.oracle_relax_tile.argtypes
        L.oracle_relax_tile_twin.restype = None
        L.tc_ip_twin.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.tc_ip_twin.restype = C.c_double
        L.oracle_det_sincos.argtypes = [_dbl_p, _dbl_p, _dbl_p, C.c_long]
        L.oracle_det_sincos.restype = None
        L.oracle_sph2cart_twin.argtypes = [_dbl_p, _dbl_p, C.c_double, C.c_long, _dbl_p]
        L.oracle_sph2cart_twin.restype = None
        L.oracle_powell_minimize_cb.argtypes = [C.c_void_p, _dbl_p, C.c_double, C.c_double, C.c_int,
                                                C.c_int, _dbl_p, _dbl_p, _int_p]
        L.oracle_powell_minimize_cb.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dbl_p)


# --------------------------------------------------------------------------- sr / es
def density_overlap_fft(rho0: np.ndarray, rho1: np.ndarray, dr) -> np.ndarray:
    """interactions.py:9-22: product of the full complex spectra of rho0 and of the
    index-reversed rho1, inverse transform scaled by the voxel volume, real part,
    then a cyclic shift by one along every axis."""
    voxel = float(np.prod(dr))
    spec = np.fft.fftn(rho0)
    spec = spec * np.fft.fftn(rho1[::-1, ::-1, ::-1])
    out = np.fft.ifftn(spec * voxel).real
    return np.roll(out, 1, axis=(0, 1, 2))


def corr_direct(rho0: np.ndarray, rho1: np.ndarray, dr, z_lo: int = 0, z_hi: int | None = None) -> np.ndarray:
    rho0 = np.ascontiguousarray(rho0, dtype=np.float64)
    rho1 = np.ascontiguousarray(rho1, dtype=np.float64)
    nx, ny, nz = rho0.shape
    z_hi = nz if z_hi is None else z_hi
    out = np.empty((nx, ny, z_hi - z_lo))
    lib().oracle_corr_direct(_p(rho0), _p(rho1), nx, ny, nz, float(np.prod(dr)), z_lo, z_hi, _p(out))
    return out


# --------------------------------------------------------------------------- tricubic
class Tricubic:
    """Stand-in for `tricubic.tricubic(list(data), list(shape))` (dbspm:562)."""

    def __init__(self, data, shape=None):
        arr = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
        if shape is not None:
            arr = arr.reshape([int(s) for s in shape])
        self._arr = arr  # keep alive: the C side borrows the buffer
        self._h = lib().tc_create(_p(arr), *[int(s) for s in arr.shape])
        if not self._h:
            raise RuntimeError("tc_create failed")

    def ip(self, xyz) -> float:
        return lib().tc_ip(self._h, float(xyz[0]), float(xyz[1]), float(xyz[2]))

    def ip_catmull(self, xyz) -> float:
        return lib().tc_ip_catmull(self._h, float(xyz[0]), float(xyz[1]), float(xyz[2]))

    def __del__(self):
        try:
            if self._h:
                lib().tc_destroy(self._h)
                self._h = None
        except Exception:
            pass


def tricubic_lm_python(data: np.ndarray, xyz) -> float:
    """Independent numpy statement of the Lekien-Marsden interpolant: solve the 64x64
    Hermite system explicitly for the voxel containing xyz (slow; tests only)."""
    n = data.shape
    d = [np.fmod(xyz[a], n[a]) for a in range(3)]
    d = [v + n[a] if v < 0 else v for a, v in enumerate(d)]
    i0 = [int(np.floor(v)) for v in d]
    u = [d[a] - i0[a] for a in range(3)]

    def at(i, j, k):
        return data[i % n[0], j % n[1], k % n[2]]

    rows, rhs = [], []
    for p in range(8):
        cx, cy, cz = p & 1, (p >> 1) & 1, (p >> 2) & 1
        a, b, c = i0[0] + cx, i0[1] + cy, i0[2] + cz
        for (ox, oy, oz) in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]:
            # central-difference estimate of d^(ox+oy+oz) f at the corner
            val = 0.0
            for sx in ([-1, 1] if ox else [0]):
                for sy in ([-1, 1] if oy else [0]):
                    for sz in ([-1, 1] if oz else [0]):
                        sign = (sx if ox else 1) * (sy if oy else 1) * (sz if oz else 1)
                        val += sign * at(a + sx, b + sy, c + sz)
            val *= 0.5 ** (ox + oy + oz)
            row = np.zeros(64)
            for k in range(4):
                for j in range(4):
                    for i in range(4):
                        def term(pw, o, x):
                            if o == 0:
                                return float(x) ** pw
                            return 0.0 if pw == 0 else pw * float(x) ** (pw - 1)
                        row[i + 4 * j + 16 * k] = term(i, ox, cx) * term(j, oy, cy) * term(k, oz, cz)
            rows.append(row)
            rhs.append(val)
    coef = np.linalg.solve(np.array(rows), np.array(rhs))
    r = 0.0
    for k in range(4):
        for j in range(4):
            for i in range(4):
                r += coef[i + 4 * j + 16 * k] * u[0] ** i * u[1] ** j * u[2] ** k
    return float(r)


# --------------------------------------------------------------------------- relax
def _emin(pol, func, origin, kappa, rpivot, dr):
    """interactions.py:55-60 with spring_energy :49-52 and sph2grid grid.py:364-369."""
    th, az = pol[0], pol[1]
    spring = 0.5 * kappa * ((np.pi - th) ** 2)
    off = [rpivot * np.sin(th) * np.cos(az) / dr[0],
           rpivot * np.sin(th) * np.sin(az) / dr[1],
           rpivot * np.cos(th) / dr[2]]
    return func(list(origin + off)) + spring


def relax_tip_z_scipy(i, j, interpE, kappa, npivot, zi, dr, method="Powell"):
    """interactions.py:70-86 on the installed scipy: int64 origin (pivot height truncates),
    top-to-bottom sweep with warm start, rows [fun, theta, phi] in ascending-z order."""
    from scipy.optimize import minimize

    origin = np.array([i, j, 0])
    start = [np.pi, 0]
    rpivot = npivot * dr[2]
    out = np.empty((len(zi), 3))
    for q in range(len(zi) - 1, -1, -1):
        origin[2] = zi[q] + npivot
        res = minimize(_emin, start, method=method, args=(interpE.ip, origin, kappa, rpivot, dr))
        start = res.x
        out[q] = [res.fun, res.x[0], res.x[1]]
    return out


def relax_tile(static_win: np.ndarray, i_lo, i_hi, j_lo, j_hi, nz_relax, kappa, rpivot, dr,
               dR=None, xtol=1e-4, ftol=1e-4, maxfev=2000, want_nfev=False, twin=False):
    """C restatement of the whole relax loop (dbspm:564-581) over a lateral tile.
    Returns (ni,nj,nz,3) [fun, theta, phi] (+ nfev).
    twin=True: the "kernel twin" mode of dbspm_oracle.c (Catmull-Rom stencil in the kernel's summation order +
    the fixed-sequence sin/cos) -- the arithmetic a GPU run must reproduce bit for bit."""
    s = np.ascontiguousarray(static_win, dtype=np.float64)
    nx, ny, nzc = s.shape
    ni, nj = i_hi - i_lo, j_hi - j_lo
    out = np.empty((ni, nj, nz_relax, 3))
    nfev = np.zeros((ni, nj, nz_relax), dtype=np.int32)
    drv = np.ascontiguousarray(dr, dtype=np.float64)
    dRp = None
    if dR is not None:
        dRv = np.ascontiguousarray(dR, dtype=np.float64).reshape(9)
        dRp = _p(dRv)
    fn = lib().oracle_relax_tile_twin if twin else lib().oracle_relax_tile
    fn(_p(s), nx, ny, nzc, i_lo, i_hi, j_lo, j_hi, nz_relax, float(kappa),
       float(rpivot), _p(drv), dRp, xtol, ftol, maxfev, _p(out), nfev.ctypes.data_as(_int_p))
    return (out, nfev) if want_nfev else out


def powell_minimize(fun, x0, xtol=1e-4, ftol=1e-4, maxiter=2000, maxfun=2000):
    """Drive the C restatement of scipy's Powell with a Python objective f(x0, x1)."""
    cb_t = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)
    cb = cb_t(lambda a, b: float(fun(np.array([a, b]))))
    x0v = np.ascontiguousarray(x0, dtype=np.float64)
    xo = np.empty(2)
    f = C.c_double()
    nit = C.c_int()
    nfev = lib().oracle_powell_minimize_cb(C.cast(cb, C.c_void_p), _p(x0v), xtol, ftol, maxiter, maxfun,
                                           _p(xo), C.byref(f), C.byref(nit))
    return xo, f.value, nit.value, nfev


def sph2cart_twin(th, az, r):
    """sph2cart with the kernel twin's sin/cos: (3, ...) array of x, y, z."""
    th = np.ascontiguousarray(th, dtype=np.float64); az = np.ascontiguousarray(az, dtype=np.float64)
    out = np.empty((3,) + th.shape)
    lib().oracle_sph2cart_twin(_p(th), _p(az), float(r), th.size, _p(out))
    return out


def det_sincos(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s, c = np.empty_like(x), np.empty_like(x)
    lib().oracle_det_sincos(_p(x), _p(s), _p(c), x.size)
    return s, c


def sph2cart(th, az, r):
    """grid.py:381-386."""
    return r * np.sin(th) * np.cos(az), r * np.sin(th) * np.sin(az), r * np.cos(th)
