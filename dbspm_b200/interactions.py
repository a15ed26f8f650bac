"""GPU versions of the reference's hot-path operators (pydbspm/interactions.py).

Same names, argument meaning and return layout as the reference so callers can switch:

  get_density_overlap(rho0, rho1, dr)                      interactions.py:9-22
  relax_tip_z(i, j, interpE, kappa, npivot, zi, dr, method) interactions.py:70-86
  relax_noortho_tip_z(..., dr, dR, method)                  interactions.py:88-105

plus the whole-grid forms the `dbspm` driver needs (one kernel launch instead of a Python
loop): density_overlap_window(), build_static(), relax_grid(), and TricubicField (the
object to pass as `interpE`; drop-in for tricubic.tricubic(list(static), list(shape))).

numpy in -> numpy out (host<->device copies inside); torch CUDA tensors in -> torch CUDA
tensors out (no copies).  All arithmetic is float64 on the GPU through the ctypes C ABI in
_lib.py; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_WS_CACHE: dict = {}


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("dbspm_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _to_dev(a, device) -> torch.Tensor:
    """float64 C-contiguous CUDA tensor from numpy / torch (host tensors go through one H2D copy)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _WS_CACHE.pop(key, None)
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


def release_workspace() -> None:
    _WS_CACHE.clear()


# ------------------------------------------------------------------------------------ sr / es
def circular_support(nonzero) -> tuple[int, int]:
    """Smallest circular interval (start, length) of plane indices that covers every True of `nonzero`
    (a tip centred on the cell origin occupies planes [0, h] and [nz-h, nz): start = nz-h, length 2h+1).
    All False -> (0, 1); no False -> (0, nz)."""
    f = np.asarray(nonzero, dtype=bool)
    nz = f.size
    if not f.any():
        return 0, 1
    if f.all():
        return 0, nz
    # longest circular run of empty planes; the support is its complement
    z = np.flatnonzero(f)
    gaps = np.diff(np.concatenate([z, [z[0] + nz]])) - 1          # empty planes after each occupied one
    g = int(np.argmax(gaps))
    start = int(z[(g + 1) % z.size])
    return start, nz - int(gaps[g])


def tip_zsupport(rho1, device=None) -> tuple[int, int]:
    """(s_lo, s_len): the z-planes (s_lo + m) mod nz, m < s_len, outside which the tip array is exactly
    zero (one streaming read of the array on the device, then a synchronising copy of nz flags)."""
    on_device = isinstance(rho1, torch.Tensor) and rho1.is_cuda
    dev = rho1.device if on_device else _device(device)
    b = _to_dev(rho1, dev)
    nx, ny, nz = (int(v) for v in b.shape)
    with torch.cuda.device(dev):
        flags = torch.empty(nz, dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().dbspm_zsupport_f64(b.data_ptr(), nx, ny, nz, flags.data_ptr(), _stream_ptr(dev)))
        return circular_support(flags.cpu().numpy() != 0)


# ---- direct real-space path for a tip with a few hundred support points ---------------------------------------
# measured on B200 (profiles/README.md, scripts/time_direct.py): the tile kernel sustains DIRECT_FMA_RATE fused
# multiply-adds per second (a quarter of the FP64 peak: 64-bit shared-memory loads of neighbouring rows conflict two-fold)
# after a padding pass that moves 16 bytes per padded sample point; the spectral path moves 80 N + 40 N_w bytes at
# SPECTRAL_BYTES_RATE (its five passes average 45 % of copy bandwidth), with N shrunk to the pruned z length for a tip
# that is compact in z.  Crossover against the pruned spectral path: K ~ 100 support points at 196x196x240, ~30 at 1024x1024x256
DIRECT_FMA_RATE = 4.0e12
SPECTRAL_BYTES_RATE = 2.9e12
DIRECT_MAX_POINTS = 4096


def tip_support_points(rho1, max_points=DIRECT_MAX_POINTS, device=None):
    """The nonzero entries of the tip array, compacted on the device: (pts (K,3) int32 indices, values (K,)) sorted by
    index, or None when there are more than max_points (a real CHGCAR tip: noise floor everywhere).  One streaming read
    of the array and a synchronising copy of at most max_points entries; do it once per tip."""
    on_device = isinstance(rho1, torch.Tensor) and rho1.is_cuda
    dev = rho1.device if on_device else _device(device)
    b = _to_dev(rho1, dev)
    nx, ny, nz = (int(v) for v in b.shape)
    with torch.cuda.device(dev):
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        idx = torch.empty(int(max_points), dtype=torch.int64, device=dev)
        val = torch.empty(int(max_points), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().dbspm_support_points_f64(b.data_ptr(), nx, ny, nz, int(max_points), count.data_ptr(),
                                                       idx.data_ptr(), val.data_ptr(), _stream_ptr(dev)))
        k = int(count.item())
    if k > int(max_points) or k == 0:
        return None
    idx_h, val_h = idx[:k].cpu().numpy(), val[:k].cpu().numpy()
    order = np.argsort(idx_h)                       # the device compaction is unordered
    idx_h, val_h = idx_h[order], val_h[order]
    pts = np.stack(np.unravel_index(idx_h, (nx, ny, nz)), axis=1).astype(np.int32)
    return pts, val_h


def _pts_args(pts):
    pts = np.ascontiguousarray(pts, dtype=np.int32)
    cols = [np.ascontiguousarray(pts[:, q]) for q in range(3)]
    return cols, [c.ctypes.data_as(C.POINTER(C.c_int)) for c in cols]


def direct_cost(shape, z_lo, z_hi, pts) -> int:
    """Fused multiply-adds per output point of the direct kernel for this support (0: it does not fit a tile)."""
    nx, ny, nz = (int(v) for v in shape)
    keep, ptrs = _pts_args(pts)
    return int(_lib.lib().dbspm_corr3d_direct_cost(nx, ny, nz, int(z_lo), int(z_hi), len(keep[0]), *ptrs))


def points_zsupport(pts, nz):
    """(s_lo, s_len) of a support given as points: what tip_zsupport finds by scanning the array."""
    flags = np.zeros(int(nz), dtype=bool)
    flags[np.asarray(pts)[:, 2]] = True
    return circular_support(flags)


def direct_is_faster(shape, z_lo, z_hi, pts) -> bool:
    """Path choice per grid shape and tip support: padding pass + N_w * (multiply-adds per output) at the measured rate of
    the tile kernel, against the bytes of the five spectral passes (on the z-pruned problem when the support allows it) at
    their measured rate."""
    nx, ny, nz = (int(v) for v in shape)
    cost = direct_cost(shape, z_lo, z_hi, pts)
    if cost == 0:
        return False
    nzw = int(z_hi) - int(z_lo)
    nw = nx * ny * nzw
    sup = points_zsupport(pts, nz)
    cut = pruned_cut(nz, z_lo, z_hi, sup) if sup[1] < nz else None
    n = nx * ny * (cut[1] if cut is not None else nz)
    zext = sup[1] - 1
    t_direct = 16.0 * nx * ny * (nzw + zext) / SPECTRAL_BYTES_RATE + nw * cost / DIRECT_FMA_RATE
    return t_direct < (80.0 * n + 40.0 * nw) / SPECTRAL_BYTES_RATE


def density_overlap_direct(rho0, pts, weights, dr, z_lo=0, z_hi=None, alpha0=1.0, out=None, device=None):
    """dV * sum_k weights[k] * rho0[(R + pts[k]) mod N]^alpha0 on the window planes: the correlation with a tip that is
    nonzero only at `pts` (tip_support_points) with weights = tip[pts]^alpha1, as a direct real-space sum."""
    on_device = isinstance(rho0, torch.Tensor) and rho0.is_cuda
    dev = rho0.device if on_device else _device(device)
    a = _to_dev(rho0, dev)
    nx, ny, nz = (int(v) for v in a.shape)
    z_hi = nz if z_hi is None else int(z_hi)
    z_lo = int(z_lo)
    keep, ptrs = _pts_args(pts)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.shape != (len(keep[0]),):
        raise ValueError("weights must have one entry per support point")
    L = _lib.lib()
    with torch.cuda.device(dev):
        need = L.dbspm_corr3d_direct_workspace_bytes(nx, ny, nz, z_lo, z_hi, len(w), *ptrs)
        if need == 0:
            raise ValueError("tip support too wide for the direct path (or bad window)")
        ws = _workspace(need, dev)
        if out is None:
            out = torch.empty((nx, ny, z_hi - z_lo), dtype=torch.float64, device=dev)
        _lib.check(L.dbspm_corr3d_direct_f64(a.data_ptr(), float(alpha0), float(np.prod(np.asarray(dr, dtype=np.float64))),
                                             nx, ny, nz, z_lo, z_hi, len(w), *ptrs, w.ctypes.data_as(C.POINTER(C.c_double)),
                                             out.data_ptr(), ws.data_ptr(), ws.numel(), 0, _stream_ptr(dev)))
    return out if on_device else out.cpu().numpy()


def density_overlap_window(rho0, rho1, dr, z_lo=0, z_hi=None, alpha0=1.0, alpha1=1.0, out=None, device=None,
                           flags=0, tip_support=None, tip_points=None):
    """dV * sum_r rho0[r]^alpha0 * rho1[(r-R) mod N]^alpha1 for R on planes z_lo..z_hi-1.

    Fuses what dbspm:329-330 / :348-349 do in three passes (`**ALPHA`, get_density_overlap,
    `[nidx]`).  Returns (Nx,Ny,z_hi-z_lo): a CUDA tensor if the inputs were CUDA tensors,
    else a numpy array.

    tip_support: None -> full-length transform; "auto" -> find the z-support of rho1 (tip_zsupport)
    and take the z-pruned path when it pays; (s_lo, s_len) -> the caller's claim that rho1 is exactly
    zero outside those planes (e.g. found once for a tip that is used for many samples).

    tip_points: None -> spectral path; "auto" -> compact the nonzero entries of rho1 (tip_support_points) and take the direct
    real-space sum when that is cheaper for this shape (direct_is_faster); (pts, values) -> the caller's list (found once
    per tip), same choice.  The path is picked per grid shape and support size."""
    if tip_points is not None:
        found = tip_support_points(rho1, device=device) if isinstance(tip_points, str) else tip_points
        if isinstance(tip_points, str) and tip_points != "auto":
            raise ValueError("tip_points must be None, 'auto' or (pts, values)")
        shape0 = tuple(int(v) for v in rho0.shape)
        zh = shape0[2] if z_hi is None else int(z_hi)
        if found is not None and direct_is_faster(shape0, z_lo, zh, found[0]):
            pts, vals = found
            with np.errstate(invalid="ignore"):
                w = vals if float(alpha1) == 1.0 else np.power(vals, float(alpha1))       # numpy's `**`, as the reference
            return density_overlap_direct(rho0, pts, w, dr, z_lo, zh, alpha0, out=out, device=device)
        if found is not None and tip_support is None and shape0[2] % 2 == 0:
            tip_support = points_zsupport(found[0], shape0[2])                            # spectral, but z-pruned: the list says where
    on_device = isinstance(rho0, torch.Tensor) and rho0.is_cuda
    dev = rho0.device if on_device else _device(device)
    a = _to_dev(rho0, dev)
    b = _to_dev(rho1, dev)
    if a.dim() != 3 or a.shape != b.shape:
        raise ValueError(f"rho0 {tuple(a.shape)} and rho1 {tuple(b.shape)} must be equal 3-D shapes")
    nx, ny, nz = (int(s) for s in a.shape)
    z_hi = nz if z_hi is None else int(z_hi)
    z_lo = int(z_lo)
    L = _lib.lib()
    if isinstance(tip_support, str):
        if tip_support != "auto":
            raise ValueError("tip_support must be None, 'auto' or (s_lo, s_len)")
        tip_support = tip_zsupport(b) if nz % 2 == 0 else None
    with torch.cuda.device(dev):
        if tip_support is None:
            need = L.dbspm_corr3d_workspace_bytes(nx, ny, nz, z_lo, z_hi, 0)
        else:
            s_lo, s_len = (int(v) for v in tip_support)
            need = L.dbspm_corr3d_pruned_workspace_bytes(nx, ny, nz, z_lo, z_hi, s_lo, s_len, 0)
        if need == 0:
            raise ValueError(f"bad window [{z_lo},{z_hi}) for shape {(nx, ny, nz)}")
        ws = _workspace(need, dev)
        if out is None:
            out = torch.empty((nx, ny, z_hi - z_lo), dtype=torch.float64, device=dev)
        elif not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
                  and tuple(out.shape) == (nx, ny, z_hi - z_lo)):
            raise ValueError("out must be a contiguous float64 CUDA tensor of the window shape")
        dV = float(np.prod(np.asarray(dr, dtype=np.float64)))
        if tip_support is None:
            _lib.check(L.dbspm_corr3d_f64(a.data_ptr(), b.data_ptr(), float(alpha0), float(alpha1), dV,
                                          nx, ny, nz, z_lo, z_hi, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                          int(flags), _stream_ptr(dev)))
        else:
            _lib.check(L.dbspm_corr3d_pruned_f64(a.data_ptr(), b.data_ptr(), float(alpha0), float(alpha1), dV,
                                                 nx, ny, nz, z_lo, z_hi, s_lo, s_len, out.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), int(flags), _stream_ptr(dev)))
    return out if on_device else out.cpu().numpy()


def get_density_overlap(rho0, rho1, dr):
    """Drop-in for pydbspm.interactions.get_density_overlap: full (Nx,Ny,Nz) energy grid."""
    return density_overlap_window(rho0, rho1, dr)


# ------------------------------------------------------------- sr / es over the ranks of a group
# x-slab decomposition (include/dbspm_b200.h, "distributed"): the three compute stages; the two exchanges
# between them (all-to-all, gather) belong to parallel.sr_es_distributed.  Buffers are float64 CUDA tensors
# whose last axis holds (re, im) pairs.
def dist_supported(shape, G: int) -> bool:
    nx, ny, nz = (int(v) for v in shape)
    return bool(_lib.lib().dbspm_corr3d_dist_supported(nx, ny, nz, int(G)))


def pruned_cut(nz, z_lo, z_hi, tip_support):
    """Geometry of the z-pruned problem for a tip that is zero outside tip_support = (s_lo, s_len): None when pruning
    does not pay (or tip_support is None), else the 7-int `cut` of include/dbspm_b200.h:
    (use, Mp, zA0, zB0, validA, validB, w_lo)."""
    if tip_support is None or int(nz) % 2:
        return None
    s_lo, s_len = (int(v) for v in tip_support)
    cut = (C.c_int * 7)()
    _lib.check(_lib.lib().dbspm_corr3d_pruned_geometry(int(nz), int(z_lo), int(z_hi), s_lo, s_len, cut))
    return tuple(int(v) for v in cut) if cut[0] else None


def _cut_arg(cut):
    return (C.c_int * 7)(*[int(v) for v in cut]) if cut is not None else None


def dist_forward(a_slab, b_slab, G, alpha0=1.0, alpha1=1.0, out=None, flags=0, cut=None):
    """y transform of this rank's x-slab (nxl,ny,nz) of both arrays, power fused, written destination-major:
    returns (sendA, sendB), each (G, nxl, ny/G, nz) float64 = complex (.., nz/2); block h goes to rank h.
    cut (pruned_cut): the pass reads the z-ranges of the pruned problem in place; nz of the buffers = cut[1]."""
    dev = a_slab.device
    a, b = _to_dev(a_slab, dev), _to_dev(b_slab, dev)
    nxl, ny, nz = (int(v) for v in a.shape)
    G = int(G)
    nzp = int(cut[1]) if cut is not None else nz
    if out is None:
        out = (torch.empty((G, nxl, ny // G, nzp), dtype=torch.float64, device=dev),
               torch.empty((G, nxl, ny // G, nzp), dtype=torch.float64, device=dev))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbspm_corr3d_dist_fwd_cut_f64(a.data_ptr(), b.data_ptr(), float(alpha0), float(alpha1), nxl, ny, nz,
                                                            G, _cut_arg(cut), out[0].data_ptr(), out[1].data_ptr(), int(flags),
                                                            _stream_ptr(dev)))
    return out


def dist_forward_peer(a_slab, b_slab, G, me, peer_ptrs, alpha0=1.0, alpha1=1.0, flags=0, cut=None):
    """dist_forward fused with the exchange: every output row is stored straight into its owner's receive buffer.
    peer_ptrs[h] = device address (int) of rank h's (2, nx, ny/G, nz) float64 receive buffer mapped into this process
    (torch symmetric memory).  The caller brackets the call with two group barriers."""
    dev = a_slab.device
    a, b = _to_dev(a_slab, dev), _to_dev(b_slab, dev)
    nxl, ny, nz = (int(v) for v in a.shape)
    ptrs = (C.c_void_p * int(G))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbspm_corr3d_dist_fwd_peer_cut_f64(a.data_ptr(), b.data_ptr(), float(alpha0), float(alpha1), nxl, ny,
                                                                 nz, int(G), int(me), _cut_arg(cut), ptrs, int(flags),
                                                                 _stream_ptr(dev)))


def dist_middle(recvA, recvB, dr, shape, G, h, z_lo, z_hi, out=None, flags=0):
    """x transform (in place) of the received spectra (nx, ny/G, nz), fused z kernel, inverse x transform:
    returns this rank's rows of the packed window, (nx, ny/G, 2*ceil(nzw/2)) float64."""
    dev = recvA.device
    nx, ny, nz = (int(v) for v in shape)
    nwh = (int(z_hi) - int(z_lo) + 1) // 2
    if out is None:
        out = torch.empty((nx, ny // int(G), 2 * nwh), dtype=torch.float64, device=dev)
    dV = float(np.prod(np.asarray(dr, dtype=np.float64)))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbspm_corr3d_dist_mid_f64(recvA.data_ptr(), recvB.data_ptr(), dV, nx, ny, nz, int(G), int(h),
                                                        int(z_lo), int(z_hi), out.data_ptr(), int(flags), _stream_ptr(dev)))
    return out


def dist_middle_peer(recvA, recvB, dr, shape, G, h, z_lo, z_hi, peer_ptrs, scratch=None, flags=0):
    """dist_middle with the second exchange fused into its inverse x pass: row x of this rank's packed window is
    stored into the (G, nx/G, ny/G, 2*ceil(nzw/2)) float64 buffer of the rank that owns x-slab x // (nx/G)
    (peer_ptrs[q] = that buffer of rank q, symmetric memory).  The caller follows with a group barrier."""
    dev = recvA.device
    nx, ny, nz = (int(v) for v in shape)
    nwh = (int(z_hi) - int(z_lo) + 1) // 2
    if scratch is None:
        scratch = torch.empty((nx, ny // int(G), 2 * nwh), dtype=torch.float64, device=dev)
    dV = float(np.prod(np.asarray(dr, dtype=np.float64)))
    ptrs = (C.c_void_p * int(G))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbspm_corr3d_dist_mid_peer_f64(recvA.data_ptr(), recvB.data_ptr(), dV, nx, ny, nz, int(G), int(h),
                                                             int(z_lo), int(z_hi), scratch.data_ptr(), ptrs, int(flags),
                                                             _stream_ptr(dev)))
    return scratch


def dist_finish(W_all, shape, G, z_lo, z_hi, out=None, flags=0):
    """inverse y transform of the gathered packed windows (G, nx, ny/G, 2*ceil(nzw/2)) -> (nx, ny, nzw)."""
    dev = W_all.device
    nx, ny, _ = (int(v) for v in shape)
    z_lo, z_hi = int(z_lo), int(z_hi)
    if out is None:
        out = torch.empty((nx, ny, z_hi - z_lo), dtype=torch.float64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        need = L.dbspm_corr3d_dist_fin_workspace_bytes(nx, ny, z_lo, z_hi)
        ws = _workspace(need, dev) if need else None
        _lib.check(L.dbspm_corr3d_dist_fin_f64(W_all.data_ptr(), nx, ny, int(G), z_lo, z_hi, out.data_ptr(),
                                               ws.data_ptr() if ws is not None else None, need, int(flags), _stream_ptr(dev)))
    return out


# ------------------------------------------------------------------------------------ static
def build_static(components, device=None):
    """static = sum_l scale_l * comp_l in the given order (dbspm:479-486: sr is scaled by V,
    the others by 1).  `components` = iterable of (array, scale).  Returns a CUDA tensor."""
    comps = list(components)
    if not comps:
        raise ValueError("no components")
    first = comps[0][0]
    dev = first.device if isinstance(first, torch.Tensor) and first.is_cuda else _device(device)
    L = _lib.lib()
    acc = None
    with torch.cuda.device(dev):
        for idx, (arr, scale) in enumerate(comps):
            t = _to_dev(arr, dev)
            if acc is None:
                acc = torch.empty_like(t)
            elif t.shape != acc.shape:
                raise ValueError("component shapes differ")
            _lib.check(L.dbspm_static_accumulate_f64(acc.data_ptr(), t.data_ptr(), float(scale), t.numel(),
                                                     1 if idx == 0 else 0, _stream_ptr(dev)))
    return acc


# ------------------------------------------------------------------------------------ relax
class TricubicField:
    """Device-resident periodic tricubic interpolant in index units.

    Drop-in for `tricubic.tricubic(list(data), list(shape))` as used at dbspm:562: `.ip([x,y,z])`
    evaluates one point (through the gather kernel); pass the object as `interpE` to
    relax_tip_z, or use relax_grid() for all lateral positions at once."""

    def __init__(self, data, shape=None, device=None):
        on_device = isinstance(data, torch.Tensor) and data.is_cuda
        dev = data.device if on_device else _device(device)
        if not isinstance(data, (torch.Tensor, np.ndarray)):
            data = np.asarray(data, dtype=np.float64)
        t = _to_dev(data, dev)
        if shape is not None:
            t = t.reshape([int(s) for s in shape])
        if t.dim() != 3:
            raise ValueError("tricubic data must be 3-D")
        self.data = t
        self.shape = tuple(int(s) for s in t.shape)

    def gather(self, pts):
        """Interpolate at pts (..., 3) in index units -> (...) values (same kind as pts)."""
        on_device = isinstance(pts, torch.Tensor) and pts.is_cuda
        dev = self.data.device
        p = _to_dev(pts, dev).reshape(-1, 3).contiguous()
        out = torch.empty(p.shape[0], dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dbspm_tricubic_gather_f64(self.data.data_ptr(), *self.shape, p.data_ptr(),
                                                            p.shape[0], out.data_ptr(), _stream_ptr(dev)))
        lead = tuple(pts.shape[:-1])
        out = out.reshape(lead)
        return out if on_device else out.cpu().numpy()

    def ip(self, xyz) -> float:
        return float(self.gather(np.asarray([xyz], dtype=np.float64))[0])


def relax_slab_halo(rpivot, dr, dR=None) -> int:
    """Planes of the static window a rank needs on each side of its x-tile (lever length in planes + the stencil)."""
    dr3 = (C.c_double * 3)(*[float(v) for v in dr])
    dR9 = None if dR is None else (C.c_double * 9)(*[float(v) for v in np.asarray(dR, dtype=np.float64).reshape(9)])
    h = int(_lib.lib().dbspm_relax_slab_halo(float(rpivot), dr3, dR9))
    if h < 0:
        raise ValueError("bad rpivot / dr / dR")
    return h


def relax_grid(static, kappa, rpivot, dr, nz_relax, dR=None, tile=None, xtol=1e-4, ftol=1e-4,
               maxfev=2000, method="Powell", want_nfev=False, device=None, coalesce=True, layout=None, x_slab=None):
    """Relax the probe at every (i, j, k) of a lateral tile (default: the whole window).

    static: (Nx,Ny,Nzc) window potential (numpy, CUDA tensor or TricubicField).
    x_slab = (x_base, Nx): `static` holds only the planes x_base, x_base + 1, ... (mod Nx) of the window -- a rank's tile
    plus relax_slab_halo() planes on both sides (multi-GPU); `tile` stays in window coordinates.  Same bits as the whole window.
    Returns dict with E (ni,nj,nz), theta, phi, tip_x, tip_y, tip_z (and nfev) -- CUDA tensors
    if `static` lives on the GPU, numpy arrays otherwise.  Equivalent to the loop at
    dbspm:564-581 followed by dbspm:597-605."""
    if str(method).lower() != "powell":
        raise NotImplementedError(f"MINMETHOD={method!r}: only scipy's Powell is restated on the GPU")
    field = static if isinstance(static, TricubicField) else None
    on_device = field is not None or (isinstance(static, torch.Tensor) and static.is_cuda)
    if field is None:
        field = TricubicField(static, device=device)
    dev = field.data.device
    nx, ny, nzc = field.shape
    nx_local, x_base = 0, 0
    if x_slab is not None:
        nx_local, (x_base, nx) = nx, (int(v) for v in x_slab)
        if tile is None:
            raise ValueError("x_slab needs an explicit tile (window coordinates)")
    i_lo, i_hi, j_lo, j_hi = (0, nx, 0, ny) if tile is None else (int(v) for v in tile)
    ni, nj, nz = i_hi - i_lo, j_hi - j_lo, int(nz_relax)
    etp = torch.empty((ni, nj, nz, 3), dtype=torch.float64, device=dev)
    cart = torch.empty((3, ni, nj, nz), dtype=torch.float64, device=dev)
    nfev = torch.empty((ni, nj, nz), dtype=torch.int32, device=dev) if want_nfev else None
    dr3 = (C.c_double * 3)(*[float(v) for v in dr])
    dR9 = None
    if dR is not None:
        dR9 = (C.c_double * 9)(*[float(v) for v in np.asarray(dR, dtype=np.float64).reshape(9)])
    if not (0 <= i_lo <= i_hi <= nx and 0 <= j_lo <= j_hi <= ny):
        raise ValueError(f"tile {(i_lo, i_hi, j_lo, j_hi)} outside the {nx}x{ny} lateral grid")
    if etp.numel() > 0:                     # an empty tile (a rank with no columns) launches nothing
        with torch.cuda.device(dev):
            # re-laid copy of the potential (include/dbspm_b200.h: y-quads, else y fastest); small tiles read in
            # place.  layout = 0 / 1 / 2 forces in-place / y fastest / y-quads (tests, A/B timing); same bits
            ws = None
            if layout is None:
                layout = -1 if (coalesce and ni * nj >= 1024) else 0
            nxs = nx_local or nx
            if layout:
                L = _lib.lib()
                need = L.dbspm_relax_workspace_bytes(nxs, ny, nzc) if layout < 0 else \
                    L.dbspm_relax_workspace_bytes_layout(nxs, ny, nzc, int(layout))
                if need:
                    ws = torch.empty(need, dtype=torch.uint8, device=dev)
            tail = (i_lo, i_hi, j_lo, j_hi, nz, float(kappa), float(rpivot),
                    dr3, dR9, float(xtol), float(ftol), int(maxfev), etp.data_ptr(), cart.data_ptr(),
                    nfev.data_ptr() if nfev is not None else None,
                    ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, _stream_ptr(dev))
            if nx_local:
                _lib.check(_lib.lib().dbspm_relax_powell_slab_f64(field.data.data_ptr(), nx, ny, nzc, x_base, nx_local, *tail))
            else:
                _lib.check(_lib.lib().dbspm_relax_powell_f64(field.data.data_ptr(), nx, ny, nzc, *tail))
    res = {"E": etp[..., 0], "theta": etp[..., 1], "phi": etp[..., 2],
           "tip_x": cart[0], "tip_y": cart[1], "tip_z": cart[2]}
    if want_nfev:
        res["nfev"] = nfev
    if not on_device:
        res = {k: v.cpu().numpy() for k, v in res.items()}
    return res


def _relax_column(i, j, interpE, kappa, npivot, zi, dr, dR, method):
    if not isinstance(interpE, TricubicField):
        raise TypeError("interpE must be a dbspm_b200.interactions.TricubicField (device-resident); "
                        "a host-side interpolator cannot be driven from the GPU and there is no CPU fallback")
    zi = np.asarray(zi)
    if zi.size == 0 or not np.array_equal(zi, np.arange(zi.size)):
        raise ValueError("zi must be arange(nz) (window-relative heights from 0, as built at dbspm:506)")
    rpivot = float(npivot) * float(dr[2])
    r = relax_grid(interpE, kappa, rpivot, dr, zi.size, dR=dR, tile=(int(i), int(i) + 1, int(j), int(j) + 1),
                   method=method)
    out = torch.stack([r["E"][0, 0], r["theta"][0, 0], r["phi"][0, 0]], dim=-1)
    return out.cpu().numpy()


def relax_tip_z(i, j, interpE, kappa, npivot, zi, dr, method="Powell"):
    """Drop-in for pydbspm.interactions.relax_tip_z: (len(zi), 3) rows [E_min, theta, phi]."""
    return _relax_column(i, j, interpE, kappa, npivot, zi, dr, None, method)


def relax_noortho_tip_z(i, j, interpE, kappa, npivot, zi, dr, dR, method="Powell"):
    """Drop-in for pydbspm.interactions.relax_noortho_tip_z."""
    return _relax_column(i, j, interpE, kappa, npivot, zi, dr, dR, method)


# ------------------------------------------------------------------------------------ next rows
def force_gradient(E, dz):
    """np.gradient(E, -dz, axis=2) on the GPU (SPMGrid.get_gradient, SPMGrid.py:375-379)."""
    on_device = isinstance(E, torch.Tensor) and E.is_cuda
    dev = E.device if on_device else _device()
    t = _to_dev(E, dev)
    out = torch.empty_like(t)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dbspm_dfz_gradient_f64(t.data_ptr(), *[int(s) for s in t.shape], float(dz),
                                                     out.data_ptr(), _stream_ptr(dev)))
    return out if on_device else out.cpu().numpy()
