"""Intra mode decision -- drop-in for IntraframeCompression/intraframe.py:24-317 on the CUDA path.

luma4x4(Y), luma16x16(Y) and chroma8x8(Cr, Cb) keep the reference's signatures and return types
(float64 planes / mode arrays); the work is done by intra.cuh through the C ABI."""
from __future__ import annotations

import numpy as np

from .runtime import get_context


def _plane(a, m, name):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8 or a.ndim != 2 or a.shape[0] % m or a.shape[1] % m:
        raise ValueError(f"{name} must be a 2-D uint8 plane with sides that are multiples of {m}")
    return a


def _run(which, m, p0, p1=None, device=0):
    H, W = p0.shape
    outs = [np.empty((H, W), np.int32) for _ in range(4 if which == 2 else 2)]
    modes = np.empty((H // m, W // m), np.uint8)
    args = [o.ctypes.data for o in outs] + [None] * (4 - len(outs))
    get_context(device).call("vcs_intra_host", which, H, W, p0.ctypes.data, p1.ctypes.data if p1 is not None else None,
                             args[0], args[1], args[2], args[3], modes.ctypes.data)
    return outs, modes


def luma4x4(Y):
    """intraframe.py:24-151 -> (Yres, Ypred, modes) as float64 arrays."""
    outs, modes = _run(0, 4, _plane(Y, 4, "Y"))
    return outs[0].astype(np.float64), outs[1].astype(np.float64), modes.astype(np.float64)


def luma16x16(Y):
    """intraframe.py:153-225 -> (Yres, Ypred, modes)."""
    outs, modes = _run(1, 16, _plane(Y, 16, "Y"))
    return outs[0].astype(np.float64), outs[1].astype(np.float64), modes.astype(np.float64)


def chroma8x8(Cr, Cb):
    """intraframe.py:228-317 -> (Crres, Crpred, Cbres, Cbpred, modes)."""
    Cr, Cb = _plane(Cr, 8, "Cr"), _plane(Cb, 8, "Cb")
    if Cr.shape != Cb.shape:
        raise ValueError("Cr and Cb must have the same shape")
    outs, modes = _run(2, 8, Cr, Cb)
    return (outs[0].astype(np.float64), outs[1].astype(np.float64), outs[2].astype(np.float64),
            outs[3].astype(np.float64), modes.astype(np.float64))
