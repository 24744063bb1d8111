"""Decoder -- drop-in for InterframeCompression/decoder.py:11-69."""
from __future__ import annotations

import numpy as np

from . import _capi
from .DCTcompressor import DCTCompressor
from .motion import WRITE_STATIC_BLOCK, MotionProcessor, _as_frame
from .runtime import get_context

WRITE_REF_FRAMES = True


def _planes_block(r, dtype):
    """The three planes as one C-contiguous [3,H,W] array without copying when they already are three consecutive
    slices of one (that is how the drop-in Encoder leaves them); otherwise one stacking copy."""
    a, b, c = (np.asarray(x) for x in r[:3])
    base = a.base
    if (base is not None and b.base is base and c.base is base and isinstance(base, np.ndarray) and base.dtype == dtype
            and base.shape == (3,) + a.shape and base.flags["C_CONTIGUOUS"]
            and a.ctypes.data == base.ctypes.data and b.ctypes.data == base[1].ctypes.data
            and c.ctypes.data == base[2].ctypes.data):
        return base
    out = np.empty((3,) + a.shape, dtype)
    out[0], out[1], out[2] = a, b, c
    return out


def _fourcc(code):
    """cv2.VideoWriter_fourcc(*code) without needing OpenCV at construction time (same little-endian packing)."""
    return sum(ord(c) << (8 * i) for i, c in enumerate(code))


class Decoder:
    def __init__(self, encoded_frames, fps, shape, ref_frames, block_size, with_DCT, dct_block_size=None):
        self.encoded_frames = encoded_frames
        self.fps = fps
        self.shape = shape
        self.fourcc = _fourcc('X264')                      # decoder.py:16
        self.ref_frames = ref_frames
        self.MotionProcessor = MotionProcessor(block_size=block_size, shape=shape)
        self.DCTCompressor = DCTCompressor(
            block_size=block_size if dct_block_size is None else dct_block_size)
        self.with_DCT = with_DCT

    def decode_frames(self, with_residuals):
        """The frames reconstruct_video would write, as a list (in-memory variant)."""
        out, num_ref_seen = [], 0
        for cur_frame in self.encoded_frames:
            if cur_frame.t == "I" and WRITE_REF_FRAMES:
                out.append(self.ref_frames[num_ref_seen])
                num_ref_seen += 1
            elif cur_frame.t == "P":
                out.append(self._reconstruct_P_frame(cur_frame, with_residuals))
        return out

    def reconstruct_video(self, with_residuals):
        """decoder.py:23-47: writes output.mp4 in the cwd with fourcc X264."""
        import cv2
        writer = cv2.VideoWriter('output.mp4', self.fourcc, self.fps,
                                 (self.shape[1], self.shape[0]))
        print("Set up video writer")
        frames = self.decode_frames(with_residuals)
        for f in frames:
            writer.write(f)
        print("Finished writing frames of length", len(self.encoded_frames) + 1)
        writer.release()
        return

    def _fully_reconstruct(self, residuals, img):
        if self.with_DCT:
            # decoder.py:55-57: img + decompress(residuals), uint8 wrap, fused into one kernel
            return self.DCTCompressor.decompress(compressed=residuals, imshape=img.shape, pred=img)
        return self.MotionProcessor._add(img, residuals)

    def _reconstruct_P_frame(self, cur_frame, with_residuals):
        ref = self.ref_frames[cur_frame.ref_i]
        if with_residuals and self.with_DCT and self.DCTCompressor.blocksize == 8:
            fused = self._reconstruct_P_frame_fused(cur_frame, ref)
            if fused is not None:
                return fused
        reconstruct_img = self.MotionProcessor.reconstruct_from_motion_vectors(
            cur_frame.mv, ref, cur_frame.c)
        if with_residuals:
            return self._fully_reconstruct(residuals=cur_frame.r, img=reconstruct_img)
        return reconstruct_img

    def _reconstruct_P_frame_fused(self, cur_frame, ref):
        """decoder.py:52-69 as ONE C-ABI call (vcs_decode_clip_host on a one-P-frame clip): motion compensation from
        the original I-frame, dequantise, IDCT, truncating store, YCrCb->BGR and the wrap add on the device."""
        mp, dc = self.MotionProcessor, self.DCTCompressor
        H, W, bs = int(mp.shape[0]), int(mp.shape[1]), int(mp.block_size)
        first = np.asarray(cur_frame.r[0])
        if first.shape != (H, W) or H % 8 or W % 8:
            return None                                    # geometry the step-by-step path handles (or refuses)
        # block_coords must be the raster grid (motion.py:74-98).  Frames of one clip normally share one coords list:
        # a list object that has been checked once is not converted and compared again.
        N = (H // bs) * (W // bs)
        if cur_frame.c is not getattr(self, "_coords_ok", None):
            want = mp._block_coords()
            coords = np.asarray(cur_frame.c, np.int64).reshape(-1, 2)
            if coords.shape != want.shape or not np.array_equal(coords, want):
                raise ValueError("block_coords must be the raster grid of _split_frame_into_mblocks")
            self._coords_ok = cur_frame.c
        mv16 = np.ascontiguousarray(np.asarray(cur_frame.mv, np.int16).reshape(-1, 2))
        if mv16.shape[0] != N:
            raise ValueError("one motion vector per macroblock expected")
        mode = _capi.COEF_I16_RINT if first.dtype == np.int16 else _capi.COEF_F64
        planes = _planes_block(cur_frame.r, np.int16 if mode == _capi.COEF_I16_RINT else np.float64)
        refc = np.ascontiguousarray(_as_frame(ref, mp.shape, "ref_frame"))
        out = np.empty((H, W, 3), np.uint8)
        ctx = get_context(mp._device)
        ctx.set_q(np.stack([np.asarray(q, np.float64) for q in dc.Q]))
        ctx.call("vcs_decode_clip_host", H, W, bs, refc.ctypes.data, 2, 2, mv16.ctypes.data, mode, planes.ctypes.data,
                 out.ctypes.data)
        num_static = int(np.count_nonzero((mv16[:, 0] == 0) & (mv16[:, 1] == 0))) if WRITE_STATIC_BLOCK else 0
        print("There are", num_static, "static blocks out of", N, "blocks")             # motion.py:67
        print("begin decompression")                                                     # DCTcompressor.py:78
        print("decompression finished")                                                  # DCTcompressor.py:90
        return out
