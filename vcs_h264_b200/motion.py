"""MotionProcessor -- drop-in for InterframeCompression/motion.py:14-161 on the CUDA path.

Same constructor, attributes, method names, argument meaning and return types as the
reference class; the work is done by libvcs_b200.so (me_generic / me_tiled kernels, mc_kernel,
wrap_kernel).  There is no NumPy fallback.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .runtime import get_context

# Optimization Params (motion.py:8-11)
SIMILARITY_THRESHOLD = 2000
CHANNELS = 3
WRITE_STATIC_BLOCK = True


def _as_frame(a, shape, name):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != CHANNELS:
        raise ValueError(f"{name} must be an HxWx3 uint8 BGR array, got {a.dtype} {a.shape}")
    if a.shape[0] < shape[0] or a.shape[1] < shape[1]:
        raise ValueError(f"{name} {a.shape} is smaller than the processor shape {shape}")
    if a.shape[0] != shape[0] or a.shape[1] != shape[1]:
        # the reference indexes with self.shape only (motion.py:82-95,128-129)
        a = np.ascontiguousarray(a[:shape[0], :shape[1]])
    return a


class MotionProcessor:
    def __init__(self, block_size, shape, device=0):
        self.block_size = block_size
        self.shape = shape
        self.search_window_size = block_size * 2          # motion.py:18
        # generalised modes (not in the reference; None keeps the reference's behaviour)
        self.step_size = None            # None -> round(block_size/3) (motion.py:132)
        self.symmetric_range = None      # R -> symmetric +/-R, clamp at H-bs (BASELINE cfg 2/3/5)
        self.metric = _capi.METRIC_WRAP8  # motion.py:146 computes the wrapped difference
        self.kernel = _capi.ME_AUTO
        self._device = device
        self.last_cost = None            # uint32[N] winning cost of the last search
        self.last_flags = None           # uint8[N]  VCS_MB_STATIC / VCS_MB_NOCAND

    # -- parameters -------------------------------------------------------------------------
    def _params(self):
        H, W = int(self.shape[0]), int(self.shape[1])
        bs = int(self.block_size)
        if self.symmetric_range is not None:
            p = _capi.me_fullsearch_params(H, W, bs, int(self.symmetric_range), self.metric,
                                           SIMILARITY_THRESHOLD)
            if self.step_size is not None:
                p.step = int(self.step_size)
        else:
            p = _capi.me_reference_params(H, W, bs)
            R = int(self.search_window_size)
            p.lo, p.hi = -R, R - bs - 1                    # motion.py:125-140
            if self.step_size is not None:
                p.step = int(self.step_size)
            p.metric = self.metric
            p.static_thr = SIMILARITY_THRESHOLD
        p.kernel = self.kernel
        return p

    # -- public API ---------------------------------------------------------------------------
    def process_motion_prediction(self, input_frame, ref_frame):
        """motion.py:20-36 -> [motion_vectors, block_coords] (lists of [dx,dy] / [x,y] ints)."""
        mv, coords = self.process_motion_prediction_arrays(input_frame, ref_frame)
        return [mv.tolist(), coords.tolist()]

    def process_motion_prediction_arrays(self, input_frame, ref_frame):
        """Array-returning variant: int32[N,2] MVs and int32[N,2] coords (no list marshalling)."""
        cur = _as_frame(input_frame, self.shape, "input_frame")
        ref = _as_frame(ref_frame, self.shape, "ref_frame")
        p = self._params()
        N = _capi.num_blocks(p.H, p.W, p.bs)
        mv = np.empty((N, 2), np.int16)
        cost = np.empty(N, np.uint32)
        flags = np.empty(N, np.uint8)
        ctx = get_context(self._device)
        ctx.call("vcs_me_search_host", p, cur.ctypes.data, ref.ctypes.data, mv.ctypes.data,
                 cost.ctypes.data, flags.ctypes.data)
        self.last_cost, self.last_flags = cost, flags
        return mv.astype(np.int32), self._block_coords()

    def get_residuals(self, input_frame, reconstructed):
        """motion.py:38-40: uint8 wrap-around difference."""
        a = np.ascontiguousarray(input_frame)
        b = np.ascontiguousarray(reconstructed)
        if a.dtype != np.uint8 or b.dtype != np.uint8 or a.shape != b.shape:
            raise ValueError("get_residuals expects two uint8 arrays of the same shape")
        out = np.empty_like(a)
        get_context(self._device).call("vcs_sub_wrap_host", a.ctypes.data, b.ctypes.data, a.size,
                                       out.ctypes.data)
        return out

    def _add(self, img, residuals):
        """Decoder._fully_reconstruct without DCT (decoder.py:60): uint8 wrap-around sum."""
        a = np.ascontiguousarray(img)
        b = np.ascontiguousarray(residuals)
        if a.dtype != np.uint8 or b.dtype != np.uint8 or a.shape != b.shape:
            raise ValueError("expects two uint8 arrays of the same shape")
        out = np.empty_like(a)
        get_context(self._device).call("vcs_add_wrap_host", a.ctypes.data, b.ctypes.data, a.size,
                                       out.ctypes.data)
        return out

    def reconstruct_from_motion_vectors(self, motion_vectors, ref_frame, block_coords):
        """motion.py:42-69: zero image + one bs x bs copy per macroblock."""
        ref = _as_frame(ref_frame, self.shape, "ref_frame")
        H, W, bs = int(self.shape[0]), int(self.shape[1]), int(self.block_size)
        mv = np.ascontiguousarray(np.asarray(motion_vectors, np.int64).reshape(-1, 2))
        coords = np.asarray(block_coords, np.int64).reshape(-1, 2)
        want = self._block_coords()
        if coords.shape != want.shape or not np.array_equal(coords, want):
            raise ValueError("block_coords must be the raster grid of _split_frame_into_mblocks")
        if mv.shape[0] != want.shape[0]:
            raise ValueError("one motion vector per macroblock expected")
        mv16 = np.ascontiguousarray(mv.astype(np.int16))
        pred = np.empty((H, W, CHANNELS), np.uint8)
        get_context(self._device).call("vcs_mc_host", H, W, bs, ref.ctypes.data, mv16.ctypes.data,
                                       pred.ctypes.data)
        num_static = int(np.count_nonzero((mv[:, 0] == 0) & (mv[:, 1] == 0))) if WRITE_STATIC_BLOCK else 0
        print("There are", num_static, "static blocks out of", len(coords), "blocks")  # motion.py:67
        return pred

    # -- private helpers kept for parity with the reference's tests/prototypes --------------------
    def _block_coords(self):
        H, W, bs = int(self.shape[0]), int(self.shape[1]), int(self.block_size)
        ys = np.arange(0, H - bs + 1, bs, dtype=np.int32)
        xs = np.arange(0, W - bs + 1, bs, dtype=np.int32)
        return np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2)   # [x,y], raster (motion.py:95)

    def _split_frame_into_mblocks(self, input_frame):
        """motion.py:74-98: views + [x,y] coords; partial blocks dropped."""
        coords = self._block_coords()
        bs = self.block_size
        blocks = [input_frame[y:y + bs, x:x + bs] for x, y in coords]
        return [blocks, coords.tolist()]

    def _get_motion_vector(self, match_coord, search_coord):
        return [match_coord[0] - search_coord[0], match_coord[1] - search_coord[1]]  # motion.py:156
