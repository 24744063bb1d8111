"""DCTCompressor -- drop-in for InterframeCompression/DCTcompressor.py:41-139 on the CUDA path.

compress()/decompress() keep the reference's signatures and return types (3 float64 planes /
BGR uint8); the arithmetic is the float64 dct_stage kernel, bit-exact with the reference.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .runtime import get_context

QF = 50.0                      # DCTcompressor.py:29
TEST_COMPRESSOR = False        # DCTcompressor.py:7
QUANTIZE = False               # DCTcompressor.py:8 (dead in the reference too)


def _default_Q():
    if "Q" not in globals():
        globals()["Q"] = list(_capi.q_tables(QF))
    return globals()["Q"]


def __getattr__(name):
    """Module attribute `Q` = [QY', QC', QC'] for (Y, Cr, Cb) (DCTcompressor.py:36-38), built on first use so
    that importing the package neither loads nor needs the CUDA library."""
    if name == "Q":
        return _default_Q()
    raise AttributeError(name)


def quality_tables(qf):
    """Q list for another quality factor (the reference edits the QF constant by hand)."""
    return list(_capi.q_tables(qf))


class DCTCompressor:
    def __init__(self, block_size, device=0):
        self.blocksize = block_size
        self.compressed = []
        self.Q = _default_Q()
        self._device = device

    def _ctx(self):
        if self.blocksize != 8:
            # the reference broadcasts an (bs,bs) block against the 8x8 tables and fails
            raise ValueError(f"operands could not be broadcast together with shapes "
                             f"({self.blocksize},{self.blocksize}) (8,8) ")
        ctx = get_context(self._device)
        ctx.set_q(np.stack([np.asarray(q, np.float64) for q in self.Q]))
        return ctx

    def compress(self, input_bgrimg, rounded=False):
        """DCTcompressor.py:49-74 -> [Y, Cr, Cb] float64 planes of D/Q (no rounding).
        rounded=True gives DCTCompression/dct.py:179's np.round(D/Q)."""
        img = np.ascontiguousarray(input_bgrimg)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("compress expects an HxWx3 uint8 BGR image")
        H, W = img.shape[:2]
        if H % 8 or W % 8:
            raise ValueError("image sides must be multiples of 8 (the reference bilinear-resizes "
                             "here, DCTcompressor.py:52; crop or resize on the host first)")
        ctx = self._ctx()
        planes = np.empty((3, H, W), np.float64)
        print("begin compression")
        ctx.call("vcs_compress_host", H, W, img.ctypes.data,
                 _capi.COEF_F64_RINT if rounded else _capi.COEF_F64, planes.ctypes.data)
        return [planes[0], planes[1], planes[2]]

    def compress_indices(self, input_bgrimg):
        """Quantised indices as int16[3,H,W] (np.round(D/Q), the compact wire format)."""
        img = np.ascontiguousarray(input_bgrimg)
        H, W = img.shape[:2]
        if img.dtype != np.uint8 or H % 8 or W % 8:
            raise ValueError("compress_indices expects an HxWx3 uint8 image, sides multiples of 8")
        ctx = self._ctx()
        idx = np.empty((3, H, W), np.int16)
        ctx.call("vcs_compress_host", H, W, img.ctypes.data, _capi.COEF_I16_RINT, idx.ctypes.data)
        return idx

    def decompress(self, compressed, imshape, pred=None):
        """DCTcompressor.py:76-93 -> BGR uint8.  pred (optional) is added mod 256
        (Decoder._fully_reconstruct, decoder.py:57)."""
        first = np.asarray(compressed[0])
        H, W = first.shape
        if H % 8 or W % 8 or imshape[0] > H or imshape[1] > W:
            raise ValueError("coefficient planes must cover imshape with sides multiples of 8")
        if first.dtype == np.int16:
            planes, mode = np.ascontiguousarray(np.stack(compressed).astype(np.int16)), _capi.COEF_I16_RINT
        else:
            planes, mode = np.ascontiguousarray(np.stack(compressed).astype(np.float64)), _capi.COEF_F64
        ctx = self._ctx()
        out = np.empty((H, W, 3), np.uint8)
        p = None
        if pred is not None:
            p = np.ascontiguousarray(pred)
            if p.dtype != np.uint8 or p.shape != out.shape:
                raise ValueError("pred must be uint8 with the planes' geometry")
        print("begin decompression")
        ctx.call("vcs_decompress_host", H, W, mode, planes.ctypes.data,
                 p.ctypes.data if p is not None else None, out.ctypes.data)
        print("decompression finished")
        return out

    # -- private helpers the reference exposes (DCTcompressor.py:100-139) -----------------------
    def _completeDCT(self, input_img):
        """DCTcompressor.py:100-109: compress then decompress (the reference then plots the result with
        matplotlib, which is the caller's business here); returns the round-tripped BGR image."""
        imshape = input_img.shape
        print("begin compression")
        compressed = self.compress(input_img)
        return self.decompress(compressed, imshape)

    def _blocks(self, matrix, inverse):
        m = np.ascontiguousarray(np.asarray(matrix, np.float64))
        if m.shape != (8, 8) or self.blocksize != 8:
            raise ValueError("only the 8x8 transform is built")
        out = np.empty((8, 8), np.float64)
        get_context(self._device).call("vcs_dct2_blocks_host", 1, int(inverse), m.ctypes.data, out.ctypes.data)
        return out

    def _dct2(self, matrix):
        """DCTcompressor.py:111-115: C . matrix . C^T (two float64 products, sequential-k FMA like np.matmul)."""
        return self._blocks(matrix, False)

    def _idct2(self, matrix):
        """DCTcompressor.py:117-121: C^T . matrix . C."""
        return self._blocks(matrix, True)

    def _dctMatrix(self):
        if self.blocksize != 8:
            raise ValueError("only the 8x8 transform is built")
        return _capi.dct_matrix()

    def _cuHelper(self, ind):
        return 2 ** -0.5 if ind == 0 else 1
