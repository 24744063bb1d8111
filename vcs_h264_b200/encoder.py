"""Encoder -- drop-in for InterframeCompression/encoder.py:9-70.

Same GOP rule: frame n is an I-frame iff n % len(pattern) == 0 (encoder.py:25), everything else
is a P-frame predicted from the ORIGINAL I-frame ref_frames[n // len(pattern)] (encoder.py:51-52).
"""
from __future__ import annotations

import math

from .DCTcompressor import DCTCompressor
from .frame import Frame
from .motion import MotionProcessor


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
        motion_vecs, coords = self.MotionProcessor.process_motion_prediction(input, ref)
        print("Finished processing motion. Got motion vectors.")
        reconstructed_img = self.MotionProcessor.reconstruct_from_motion_vectors(motion_vecs, ref, coords)
        residuals = self.MotionProcessor.get_residuals(input_frame=input, reconstructed=reconstructed_img)
        res = self.DCTCompressor.compress(residuals) if self.with_DCT else residuals
        return Frame("P", motion_vectors=motion_vecs, residuals=res, block_coords=coords,
                     index=frame_num, ref_idx=ref_idx)
