"""Pinned-memory reader for the reference's `.npz` files.

`np.savez` (grid.py:165-176, vasp.py writers) stores every array as an uncompressed `.npy`
member of a zip file.  `np.load` reads a member through `zipfile`, which copies it through a
CRC-checking stream into pageable memory; the host->device copy that follows is staged a second
time by the driver.  For the 2 GB arrays of the large configurations that is most of the
wall-clock of an sr/es step.  Here the member's byte range is located from the zip directory,
the `.npy` header parsed, and the payload read with `readinto` straight into a pinned torch
tensor, which is then copied to the device asynchronously (read of the next array overlaps
the copy of the previous one).  Compressed members, object arrays and Fortran-ordered arrays
fall back to `np.load` semantics (returned as ordinary numpy arrays)."""
from __future__ import annotations

import struct
import zipfile

import numpy as np
import torch
from numpy.lib import format as npy_format

_LOCAL_HEADER = struct.Struct("<4s5H3L2H")      # zip local file header (30 bytes)


def _member_payload(fh, info: zipfile.ZipInfo):
    """-> (offset of the raw .npy bytes, length) of a stored (uncompressed) zip member."""
    fh.seek(info.header_offset)
    raw = fh.read(_LOCAL_HEADER.size)
    sig, _, _, method, _, _, _, _, _, n_name, n_extra = _LOCAL_HEADER.unpack(raw)
    if sig != b"PK\x03\x04":
        raise ValueError("bad zip local header")
    return info.header_offset + _LOCAL_HEADER.size + n_name + n_extra, info.file_size, method


def read_npz(path, keys=None, pin=True, device=None, dtype=None):
    """Read arrays of an `.npz`.

    keys: names to read (default: all).  pin: read numeric C-ordered members into pinned memory.
    device: if given, such members are also copied to that CUDA device (non-blocking, overlapped
    with the reading of the next member) and returned as CUDA tensors; everything else comes back
    as numpy arrays (small metadata: cell, positions, numbers ...).  dtype: optional torch dtype
    the device copies are converted to (e.g. torch.float64)."""
    out = {}
    pin = pin and torch.cuda.is_available()
    with zipfile.ZipFile(path) as zf, open(path, "rb") as fh:
        names = {i.filename[:-4] if i.filename.endswith(".npy") else i.filename: i for i in zf.infolist()}
        for key in (keys if keys is not None else names):
            if key not in names:
                raise KeyError(f"{key} is not a member of {path}")
            info = names[key]
            off, size, method = _member_payload(fh, info)
            direct = method == zipfile.ZIP_STORED
            if direct:
                fh.seek(off)
                version = npy_format.read_magic(fh)
                shape, fortran, dt = (npy_format.read_array_header_1_0(fh) if version == (1, 0)
                                      else npy_format.read_array_header_2_0(fh))
                direct = not (dt.hasobject or fortran) and dt.kind in "fiub" and dt.isnative
            if not direct or int(np.prod(shape)) * dt.itemsize < (1 << 20):
                with zf.open(info) as mf:
                    out[key] = npy_format.read_array(mf, allow_pickle=True)
                continue
            t = torch.empty(tuple(shape), dtype=torch.from_numpy(np.empty(0, dtype=dt)).dtype, pin_memory=pin)
            buf = memoryview(t.numpy().reshape(-1).view(np.uint8))
            got, chunk = 0, 64 << 20
            while got < len(buf):
                n = fh.readinto(buf[got:got + chunk])
                if not n:
                    raise EOFError(f"{path}:{key} truncated")
                got += n
            if device is not None:
                d = t.to(device, non_blocking=True)
                out[key] = d.to(dtype) if dtype is not None and d.dtype != dtype else d
            else:
                out[key] = t
    if device is not None and torch.cuda.is_available():
        torch.cuda.current_stream(device).synchronize()      # pinned staging buffers may now be released
    return out
