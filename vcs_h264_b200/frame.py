"""Frame record -- same attribute bag as InterframeCompression/frame.py:1-8."""


class Frame:
    def __init__(self, frame_type, motion_vectors, residuals, block_coords, index, ref_idx):
        self.t, self.mv, self.r = frame_type, motion_vectors, residuals   # "I"/"P", MVs, residual
        self.c, self.i, self.ref_i = block_coords, index, ref_idx
