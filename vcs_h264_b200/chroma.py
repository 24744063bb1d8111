"""4:2:0 chroma subsampling -- the arithmetic of ChromaSubsampling/chroma.py on the CUDA path.

The reference is a script without functions; its two stages are exposed here under the names of its own
variables:  subsample420(img) -> [Y, crSamples, cbSamples]  (chroma.py:9-24) and
reconstruct(subsampledImg) -> finalImg  (chroma.py:27-41).  The work is done by chroma.cuh through the C ABI."""
from __future__ import annotations

import numpy as np

from .runtime import get_context


def _bgr(img):
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3 or 0 in img.shape:
        raise ValueError("img must be a non-empty H x W x 3 uint8 BGR image")
    return img


def subsample420(img, device=0, with_reconstruction=False):
    """chroma.py:9-24: [Y (H x W), crSamples, cbSamples (ceil(H/2) x ceil(W/2))], all uint8.
    with_reconstruction=True also returns chroma.py's finalImg from the same device pass."""
    img = _bgr(img)
    H, W, _ = img.shape
    Y = np.empty((H, W), np.uint8)
    cr = np.empty(((H + 1) // 2, (W + 1) // 2), np.uint8)
    cb = np.empty_like(cr)
    final = np.empty_like(img) if with_reconstruction else None
    get_context(device).call("vcs_chroma420_host", H, W, img.ctypes.data, Y.ctypes.data, cr.ctypes.data, cb.ctypes.data,
                             final.ctypes.data if final is not None else None)
    return ([Y, cr, cb], final) if with_reconstruction else [Y, cr, cb]


def reconstruct(subsampledImg, device=0):
    """chroma.py:27-41: finalImg (H x W x 3 uint8 BGR) from [Y, crSamples, cbSamples]."""
    Y, cr, cb = (np.ascontiguousarray(p) for p in subsampledImg)
    if Y.ndim != 2 or 0 in Y.shape:
        raise ValueError("Y must be a non-empty 2-D plane")
    H, W = Y.shape
    if Y.dtype != np.uint8 or cr.dtype != np.uint8 or cb.dtype != np.uint8 or \
            cr.shape != ((H + 1) // 2, (W + 1) // 2) or cb.shape != cr.shape:
        raise ValueError("subsampledImg must be [Y (H x W), Cr, Cb (ceil(H/2) x ceil(W/2))] uint8 planes")
    out = np.empty((H, W, 3), np.uint8)
    get_context(device).call("vcs_chroma420_to_bgr_host", H, W, Y.ctypes.data, cr.ctypes.data, cb.ctypes.data,
                             out.ctypes.data)
    return out
