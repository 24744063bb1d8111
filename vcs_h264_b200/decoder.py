"""Decoder -- drop-in for InterframeCompression/decoder.py:11-69."""
from __future__ import annotations

from .DCTcompressor import DCTCompressor
from .motion import MotionProcessor

WRITE_REF_FRAMES = True


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
        reconstruct_img = self.MotionProcessor.reconstruct_from_motion_vectors(
            cur_frame.mv, ref, cur_frame.c)
        if with_residuals:
            return self._fully_reconstruct(residuals=cur_frame.r, img=reconstruct_img)
        return reconstruct_img
