"""ctypes front-end of oracle/vcs_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (as the checker / the timed CPU port).  The product package
vcs_h264_b200 never does.

Every wrapper cites the reference function it restates (file:line under /root/reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvcs_oracle.so")

METRIC_WRAP8 = 0
METRIC_SAD = 1


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a few seconds)."""
    src = os.path.join(_HERE, "vcs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libvcs_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.vcs_oracle_me.restype = C.c_int
        _lib.vcs_oracle_encode_p.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _u8(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8
    return a


def reference_search_params(bs: int, R: int | None = None, step: int | None = None):
    """(lo, hi, step, slack) of MotionProcessor._find_match's own loop
    (InterframeCompression/motion.py:18,123-140): R = 2*bs, step = round(bs/3)."""
    R = 2 * bs if R is None else R
    step = round(bs / 3) if step is None else step
    return dict(lo=-R, hi=R - bs - 1, step=step, slack=1)


def symmetric_search_params(R: int):
    """(lo, hi, step, slack) of the +/-R step-1 full search of BASELINE configs 2/3/5."""
    return dict(lo=-R, hi=R, step=1, slack=0)


def me(cur, ref, bs, lo, hi, step, slack, metric=METRIC_WRAP8, static_thr=2000, simd=True,
       nthreads=0):
    """MotionProcessor.process_motion_prediction (motion.py:20-36).
    Returns mv int32[N,2] ([dx,dy]), cost uint32[N], flags uint8[N] (1 static, 2 none)."""
    cur, ref = _u8(cur), _u8(ref)
    H, W, _ = cur.shape
    N = (H // bs) * (W // bs)
    mv = np.zeros((N, 2), np.int32)
    cost = np.zeros(N, np.uint32)
    flags = np.zeros(N, np.uint8)
    rc = lib().vcs_oracle_me(_p(cur, C.c_uint8), _p(ref, C.c_uint8), H, W, bs, lo, hi, step, slack,
                             metric, C.c_longlong(static_thr), int(simd), nthreads,
                             _p(mv, C.c_int32), _p(cost, C.c_uint32), _p(flags, C.c_uint8))
    if rc:
        raise ValueError(f"vcs_oracle_me rc={rc}")
    return mv, cost, flags


def block_coords(H, W, bs):
    """MotionProcessor._split_frame_into_mblocks coords (motion.py:74-98): [x,y] raster."""
    return np.array([[x, y] for y in range(0, H - bs + 1, bs) for x in range(0, W - bs + 1, bs)],
                    np.int32).reshape(-1, 2)


def mc(ref, bs, mv):
    """MotionProcessor.reconstruct_from_motion_vectors (motion.py:42-69)."""
    ref = _u8(ref)
    H, W, _ = ref.shape
    mv = np.ascontiguousarray(mv, np.int32)
    pred = np.empty_like(ref)
    rc = lib().vcs_oracle_mc(_p(ref, C.c_uint8), H, W, bs, _p(mv, C.c_int32), _p(pred, C.c_uint8))
    if rc:
        raise ValueError(f"vcs_oracle_mc rc={rc}")
    return pred


def residual(cur, pred):
    """MotionProcessor.get_residuals (motion.py:38-40)."""
    cur, pred = _u8(cur), _u8(pred)
    out = np.empty_like(cur)
    lib().vcs_oracle_residual(_p(cur, C.c_uint8), _p(pred, C.c_uint8), C.c_size_t(cur.size),
                              _p(out, C.c_uint8))
    return out


def add_wrap(a, b):
    """Decoder._fully_reconstruct (decoder.py:57)."""
    a, b = _u8(a), _u8(b)
    out = np.empty_like(a)
    lib().vcs_oracle_add_wrap(_p(a, C.c_uint8), _p(b, C.c_uint8), C.c_size_t(a.size),
                              _p(out, C.c_uint8))
    return out


def bgr2ycrcb(img):
    """cv2.cvtColor(img, COLOR_BGR2YCR_CB) (DCTcompressor.py:55)."""
    img = _u8(img)
    out = np.empty_like(img)
    lib().vcs_oracle_bgr2ycrcb(_p(img, C.c_uint8), C.c_size_t(img.size // 3), _p(out, C.c_uint8))
    return out


def ycrcb2bgr(img):
    """cv2.cvtColor(img, COLOR_YCR_CB2BGR) (DCTcompressor.py:92)."""
    img = _u8(img)
    out = np.empty_like(img)
    lib().vcs_oracle_ycrcb2bgr(_p(img, C.c_uint8), C.c_size_t(img.size // 3), _p(out, C.c_uint8))
    return out


def dct_matrix():
    """DCTCompressor._dctMatrix (DCTcompressor.py:124-133)."""
    m = np.empty((8, 8), np.float64)
    lib().vcs_oracle_dct_matrix(_p(m, C.c_double))
    return m


def qtables(qf=50.0):
    """Q list of DCTcompressor.py:29-38 / dct.py:157-166 as float64[3,8,8] (Y,Cr,Cb)."""
    q = np.empty((3, 8, 8), np.float64)
    rc = lib().vcs_oracle_qtables(C.c_double(qf), _p(q, C.c_double))
    if rc:
        raise ValueError("Invalid quality setting, must be between 1 and 99.")
    return q


def dct2(block):
    """DCTCompressor._dct2 (DCTcompressor.py:111-115)."""
    x = np.ascontiguousarray(block, np.float64)
    d = np.empty((8, 8), np.float64)
    lib().vcs_oracle_dct2(_p(x, C.c_double), _p(d, C.c_double))
    return d


def idct2(block):
    """DCTCompressor._idct2 (DCTcompressor.py:117-121)."""
    x = np.ascontiguousarray(block, np.float64)
    d = np.empty((8, 8), np.float64)
    lib().vcs_oracle_idct2(_p(x, C.c_double), _p(d, C.c_double))
    return d


def compress(bgr, Q=None, round_mode=0, nthreads=0):
    """DCTCompressor.compress (DCTcompressor.py:49-74); round_mode=1 is dct.py:179.
    Returns float64[3,H,W] (Y,Cr,Cb planes)."""
    bgr = _u8(bgr)
    H, W, _ = bgr.shape
    Q = qtables(50.0) if Q is None else np.ascontiguousarray(Q, np.float64)
    planes = np.empty((3, H, W), np.float64)
    rc = lib().vcs_oracle_compress(_p(bgr, C.c_uint8), H, W, _p(Q, C.c_double), round_mode,
                                   nthreads, _p(planes, C.c_double))
    if rc:
        raise ValueError(f"vcs_oracle_compress rc={rc} (H, W must be multiples of 8)")
    return planes


def decompress(planes, Q=None, nthreads=0):
    """DCTCompressor.decompress (DCTcompressor.py:76-93)."""
    planes = np.ascontiguousarray(planes, np.float64)
    _, H, W = planes.shape
    Q = qtables(50.0) if Q is None else np.ascontiguousarray(Q, np.float64)
    bgr = np.empty((H, W, 3), np.uint8)
    rc = lib().vcs_oracle_decompress(_p(planes, C.c_double), H, W, _p(Q, C.c_double), nthreads,
                                     _p(bgr, C.c_uint8))
    if rc:
        raise ValueError(f"vcs_oracle_decompress rc={rc}")
    return bgr


def encode_p(cur, ref, bs, lo, hi, step, slack, metric=METRIC_WRAP8, static_thr=2000, Q=None,
             round_mode=0, simd=True, nthreads=0, want_planes=True, want_recon=True):
    """Encoder._process_P_frame + Decoder._reconstruct_P_frame (encoder.py:49-70,
    decoder.py:52-69) for one P-frame.  Returns dict(mv, cost, flags, planes, recon)."""
    cur, ref = _u8(cur), _u8(ref)
    H, W, _ = cur.shape
    N = (H // bs) * (W // bs)
    Q = qtables(50.0) if Q is None else np.ascontiguousarray(Q, np.float64)
    mv = np.zeros((N, 2), np.int32)
    cost = np.zeros(N, np.uint32)
    flags = np.zeros(N, np.uint8)
    planes = np.empty((3, H, W), np.float64) if want_planes else None
    recon = np.empty((H, W, 3), np.uint8) if want_recon else None
    rc = lib().vcs_oracle_encode_p(_p(cur, C.c_uint8), _p(ref, C.c_uint8), H, W, bs, lo, hi, step,
                                   slack, metric, C.c_longlong(static_thr), _p(Q, C.c_double),
                                   round_mode, int(simd), nthreads, _p(mv, C.c_int32),
                                   _p(cost, C.c_uint32), _p(flags, C.c_uint8),
                                   _p(planes, C.c_double), _p(recon, C.c_uint8))
    if rc:
        raise ValueError(f"vcs_oracle_encode_p rc={rc}")
    return dict(mv=mv, cost=cost, flags=flags, planes=planes, recon=recon)


class ForwardEncoder:
    """bench.py's CPU arm: Encoder._process_P_frame (encoder.py:49-70) with the rounded quantiser, forward half
    only, int8 indices -- the same outputs as the GPU end-to-end leg -- into buffers allocated ONCE here."""

    def __init__(self, H, W, bs, nP):
        N = (H // bs) * (W // bs)
        self.H, self.W, self.bs = H, W, bs
        self.mv = np.zeros((nP, N, 2), np.int32)
        self.cost = np.zeros((nP, N), np.uint32)
        self.flags = np.zeros((nP, N), np.uint8)
        self.coef = np.zeros((nP, 3, H, W), np.int8)
        self.scratch = np.zeros(2 * H * W * 3, np.uint8)

    def encode(self, p, cur, ref, lo, hi, step, slack, metric=METRIC_WRAP8, static_thr=2000, Q=None, simd=True,
               nthreads=0):
        cur, ref = _u8(cur), _u8(ref)
        Q = qtables(50.0) if Q is None else Q
        rc = lib().vcs_oracle_encode_p_i8(_p(cur, C.c_uint8), _p(ref, C.c_uint8), self.H, self.W, self.bs, lo, hi,
                                          step, slack, metric, C.c_longlong(static_thr), _p(Q, C.c_double),
                                          int(simd), nthreads, _p(self.mv[p], C.c_int32), _p(self.cost[p], C.c_uint32),
                                          _p(self.flags[p], C.c_uint8), _p(self.coef[p], C.c_int8),
                                          _p(self.scratch, C.c_uint8))
        if rc:
            raise ValueError(f"vcs_oracle_encode_p_i8 rc={rc}")


def max_threads():
    return lib().vcs_oracle_max_threads()


def selfcheck_simd():
    return lib().vcs_oracle_selfcheck_simd()


def _plane(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8 and a.ndim == 2
    return a


def luma4x4(Y):
    """luma4x4 (IntraframeCompression/intraframe.py:24-151) -> (res int32, pred int32, modes uint8)."""
    Y = _plane(Y)
    H, W = Y.shape
    res = np.empty((H, W), np.int32); pred = np.empty((H, W), np.int32); modes = np.empty((H // 4, W // 4), np.uint8)
    rc = lib().vcs_oracle_luma4x4(_p(Y, C.c_uint8), H, W, _p(res, C.c_int32), _p(pred, C.c_int32), _p(modes, C.c_uint8))
    if rc:
        raise ValueError("luma4x4 needs sides that are multiples of 4")
    return res, pred, modes


def luma16x16(Y):
    """luma16x16 (intraframe.py:153-225)."""
    Y = _plane(Y)
    H, W = Y.shape
    res = np.empty((H, W), np.int32); pred = np.empty((H, W), np.int32); modes = np.empty((H // 16, W // 16), np.uint8)
    rc = lib().vcs_oracle_luma16x16(_p(Y, C.c_uint8), H, W, _p(res, C.c_int32), _p(pred, C.c_int32), _p(modes, C.c_uint8))
    if rc:
        raise ValueError("luma16x16 needs sides that are multiples of 16")
    return res, pred, modes


def chroma8x8(Cr, Cb):
    """chroma8x8 (intraframe.py:228-317) -> (Crres, Crpred, Cbres, Cbpred int32, modes uint8)."""
    Cr, Cb = _plane(Cr), _plane(Cb)
    H, W = Cr.shape
    outs = [np.empty((H, W), np.int32) for _ in range(4)]
    modes = np.empty((H // 8, W // 8), np.uint8)
    rc = lib().vcs_oracle_chroma8x8(_p(Cr, C.c_uint8), _p(Cb, C.c_uint8), H, W, *[_p(o, C.c_int32) for o in outs],
                                    _p(modes, C.c_uint8))
    if rc:
        raise ValueError("chroma8x8 needs sides that are multiples of 8")
    return (*outs, modes)


def chroma420(bgr):
    """ChromaSubsampling/chroma.py:9-21: BGR -> [Y (H x W), crSamples, cbSamples (ceil(H/2) x ceil(W/2))]."""
    bgr = _u8(bgr)
    H, W, _ = bgr.shape
    Y = np.empty((H, W), np.uint8)
    cr = np.empty(((H + 1) // 2, (W + 1) // 2), np.uint8)
    cb = np.empty_like(cr)
    rc = lib().vcs_oracle_chroma420(_p(bgr, C.c_uint8), H, W, _p(Y, C.c_uint8), _p(cr, C.c_uint8), _p(cb, C.c_uint8))
    if rc:
        raise ValueError(f"vcs_oracle_chroma420 rc={rc}")
    return Y, cr, cb


def chroma420_to_bgr(Y, cr, cb):
    """ChromaSubsampling/chroma.py:27-41 (NumPy-2 uint8 scalar wrap of `Cr - 128`, float64, truncating store)."""
    Y, cr, cb = _plane(Y), _plane(cr), _plane(cb)
    H, W = Y.shape
    assert cr.shape == ((H + 1) // 2, (W + 1) // 2) == cb.shape
    out = np.empty((H, W, 3), np.uint8)
    rc = lib().vcs_oracle_chroma420_to_bgr(_p(Y, C.c_uint8), _p(cr, C.c_uint8), _p(cb, C.c_uint8), H, W, _p(out, C.c_uint8))
    if rc:
        raise ValueError(f"vcs_oracle_chroma420_to_bgr rc={rc}")
    return out
