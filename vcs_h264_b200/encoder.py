"""Encoder -- drop-in for InterframeCompression/encoder.py:9-70.

Same GOP rule: frame n is an I-frame iff n % len(pattern) == 0 (encoder.py:25), everything else
is a P-frame predicted from the ORIGINAL I-frame ref_frames[n // len(pattern)] (encoder.py:51-52).
"""
from __future__ import annotations

import math

import numpy as np

from . import _capi
from .DCTcompressor import DCTCompressor
from .frame import Frame
from .motion import WRITE_STATIC_BLOCK, MotionProcessor, _as_frame
from .runtime import get_context


class Encoder:
    def __init__(self, pattern, shape, block_size, with_DCT, dct_block_size=None):
        self.ref_frames = []
        self.encoded_frames = []
        self.pattern = pattern
        self.ENCODING_PATTERN_LENGTH = len(pattern)
        self.MotionProcessor = MotionProcessor(block_size=block_size, shape=shape)
        # The reference hands the ME block size to the DCT too (encoder.py:18), which only works
        # for 8.  dct_block_size=8 decouples them (BASELINE config 2: 16x16 ME + 8x8 DCT).
        self.DCTCompressor = DCTCompressor(
            block_size=block_size if dct_block_size is None else dct_block_size)
        self.with_DCT = with_DCT
        self._pair = None          # [ref, cur] staging of the fused P-frame call
        self._coords = None        # raster [x, y] list, built once (every frame has the same grid)

    def encode_frame(self, input_frame, frame_num):
        print("Encoding new frame of index", frame_num)
        if frame_num % self.ENCODING_PATTERN_LENGTH == 0:
            encoded_frame, frame_type = self._process_I_frame(input_frame, frame_num), "I"
        else:
            encoded_frame, frame_type = self._process_P_frame(input_frame, frame_num), "P"
        print("Encoded frame of type", frame_type)
        self.encoded_frames.append(encoded_frame)
        return

    def _process_I_frame(self, input_frame, frame_num):
        self.ref_frames.append(input_frame)                       # original frame (encoder.py:42)
        return Frame("I", None, None, None, frame_num, frame_num % self.ENCODING_PATTERN_LENGTH)

    def _process_B_frame(self, input, frame_num):
        return                                                     # stub in the reference too (:45-47)

    def _process_P_frame(self, input, frame_num):
        print("Processing P frame")
        ref_idx = math.floor(frame_num / self.ENCODING_PATTERN_LENGTH)
        ref = self.ref_frames[ref_idx]
        if self.with_DCT and self.DCTCompressor.blocksize == 8 and not (input.shape[0] % 8 or input.shape[1] % 8):
            return self._process_P_frame_fused(input, ref, frame_num, ref_idx)
        motion_vecs, coords = self.MotionProcessor.process_motion_prediction(input, ref)
        print("Finished processing motion. Got motion vectors.")
        reconstructed_img = self.MotionProcessor.reconstruct_from_motion_vectors(motion_vecs, ref, coords)
        residuals = self.MotionProcessor.get_residuals(input_frame=input, reconstructed=reconstructed_img)
        res = self.DCTCompressor.compress(residuals) if self.with_DCT else residuals
        return Frame("P", motion_vectors=motion_vecs, residuals=res, block_coords=coords,
                     index=frame_num, ref_idx=ref_idx)

    def _process_P_frame_fused(self, input, ref, frame_num, ref_idx):
        """encoder.py:49-70 as ONE C-ABI call (vcs_encode_clip_host on the 2-frame clip [ref, input]): search,
        motion compensation, wrap residual and the float64 DCT/quantiser run back to back on the device and only
        the vectors and the three coefficient planes come back.  Same objects, same prints as the step-by-step path;
        it honours whatever was set on self.MotionProcessor (window, step, metric) and self.DCTCompressor.Q."""
        mp, dc = self.MotionProcessor, self.DCTCompressor
        H, W = int(mp.shape[0]), int(mp.shape[1])
        pair = self._pair if getattr(self, "_pair", None) is not None and self._pair.shape == (2, H, W, 3) else None
        if pair is None:
            pair = self._pair = np.empty((2, H, W, 3), np.uint8)
        pair[0] = _as_frame(ref, mp.shape, "ref_frame")
        pair[1] = _as_frame(input, mp.shape, "input_frame")
        p = mp._params()
        N = _capi.num_blocks(p.H, p.W, p.bs)
        mv = np.empty((N, 2), np.int16)
        cost = np.empty(N, np.uint32)
        flags = np.empty(N, np.uint8)
        planes = np.empty((3, H, W), np.float64)
        ctx = get_context(mp._device)
        ctx.set_q(np.stack([np.asarray(q, np.float64) for q in dc.Q]))
        ctx.call("vcs_encode_clip_host", p, pair.ctypes.data, 2, 2, _capi.COEF_F64, mv.ctypes.data, cost.ctypes.data,
                 flags.ctypes.data, planes.ctypes.data, None)
        mp.last_cost, mp.last_flags = cost, flags
        if self._coords is None or len(self._coords) != N:
            self._coords = mp._block_coords().tolist()
        motion_vecs = mv.astype(np.int32).tolist()
        print("Finished processing motion. Got motion vectors.")
        num_static = int(np.count_nonzero((mv[:, 0] == 0) & (mv[:, 1] == 0))) if WRITE_STATIC_BLOCK else 0
        print("There are", num_static, "static blocks out of", N, "blocks")          # motion.py:67
        print("begin compression")                                                     # DCTcompressor.py:61
        return Frame("P", motion_vectors=motion_vecs, residuals=[planes[0], planes[1], planes[2]],
                     block_coords=self._coords, index=frame_num, ref_idx=ref_idx)
