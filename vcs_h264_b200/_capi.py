"""ctypes binding of libvcs_b200.so (include/vcs_b200.h).

Loading fails loudly when the CUDA library is missing or cannot be built -- there is no CPU
fallback and nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvcs_b200.so")

OK = 0
METRIC_WRAP8, METRIC_SAD = 0, 1
MB_STATIC, MB_NOCAND = 1, 2
COEF_F64, COEF_F64_RINT, COEF_I16_RINT, COEF_I8_RINT = 0, 1, 2, 3
ME_AUTO, ME_GENERIC, ME_TILED = 0, 1, 2


class MeParams(C.Structure):
    """struct vcs_me_params (include/vcs_b200.h)."""
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("bs", C.c_int32), ("lo", C.c_int32),
                ("hi", C.c_int32), ("step", C.c_int32), ("slack", C.c_int32),
                ("metric", C.c_int32), ("static_thr", C.c_int64), ("kernel", C.c_int32),
                ("reserved", C.c_int32)]

    def copy(self):
        q = MeParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(MeParams))
        return q


class VcsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vcs_b200 error {code}: {msg}")
        self.code = code


_vp, _i, _sz, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_double
_PP = C.POINTER(MeParams)

# name -> (restype, argtypes); every symbol include/vcs_b200.h declares
SIGNATURES = {
    "vcs_version": (_i, []),
    "vcs_create": (_i, [_i, C.POINTER(_vp)]),
    "vcs_destroy": (_i, [_vp]),
    "vcs_last_error": (C.c_char_p, [_vp]),
    "vcs_set_stream": (_i, [_vp, _vp]),
    "vcs_use_own_stream": (_i, [_vp]),
    "vcs_synchronize": (_i, [_vp]),
    "vcs_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    "vcs_launch_count": (C.c_int64, [_vp]),
    "vcs_me_reference_params": (_i, [_i, _i, _i, _PP]),
    "vcs_me_fullsearch_params": (_i, [_i, _i, _i, _i, _i, C.c_int64, _PP]),
    "vcs_num_blocks": (_i, [_i, _i, _i]),
    "vcs_q_tables": (_i, [_d, _vp]),
    "vcs_dct_matrix": (_i, [_vp]),
    "vcs_set_q": (_i, [_vp, _vp]),
    "vcs_me_search_dev": (_i, [_vp, _PP, _vp, _vp, _vp, _vp, _vp]),
    "vcs_me_search_host": (_i, [_vp, _PP, _vp, _vp, _vp, _vp, _vp]),
    "vcs_num_p_frames": (_i, [_i, _i]),
    "vcs_me_search_clip_dev": (_i, [_vp, _PP, _vp, _i, _i, _vp, _vp, _vp]),
    "vcs_mc_dev": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "vcs_mc_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "vcs_sub_wrap_dev": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "vcs_add_wrap_dev": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "vcs_sub_wrap_host": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "vcs_add_wrap_host": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "vcs_dct2_blocks_host": (_i, [_vp, _i, _i, _vp, _vp]),
    "vcs_compress_dev": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "vcs_compress_host": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "vcs_decompress_dev": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "vcs_decompress_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "vcs_residual_dct_clip_dev": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "vcs_encode_clip_dev": (_i, [_vp, _PP, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vcs_encode_clip_host": (_i, [_vp, _PP, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vcs_encode_clip_host_packed": (_i, [_vp, _PP, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz,
                                         C.POINTER(C.c_uint64), _vp]),
    "vcs_pack_coef_dev": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint64)]),
    "vcs_unpack_coef_dev": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp]),
    "vcs_decode_clip_host_packed": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp]),
    "vcs_decode_clip_dev": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "vcs_decode_clip_host": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp]),
    "vcs_count_nonzero_dev": (_i, [_vp, _i, _vp, _sz, C.POINTER(C.c_ulonglong)]),
    "vcs_set_dct_precision": (_i, [_vp, _i]),
    "vcs_flip_counters_dev": (_i, [_vp, _i, _vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(C.c_ulonglong)]),
    "vcs_intra_luma4x4_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vcs_intra_luma16x16_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vcs_intra_chroma8x8_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vcs_intra_host": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vcs_chroma420_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vcs_chroma420_to_bgr_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vcs_chroma420_host": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vcs_chroma420_to_bgr_host": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "vcs_microbench": (_i, [_vp, _i, _i, C.POINTER(_d), C.POINTER(_d)]),
    "vcs_enable_kernel_timing": (_i, [_vp, _i]),
    "vcs_kernel_times": (_i, [_vp, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i)]),
}

_lib = None


def load():
    """dlopen the in-tree library.  Never compiles: `python -m vcs_h264_b200.build` (or
    __graft_entry__.build()) does that; a missing library is an error, a stale one a warning."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("VCS_B200_LIB")   # A/B builds of the same sources (tools/ab_me.sh); normally unset
    if not path:
        path = LIB_PATH
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is missing: build it with `python -m vcs_h264_b200.build` "
                          "(nvcc, sm_100a); this package has no CPU path")
        from . import build as _build
        if _build.stale():
            import warnings
            warnings.warn("libvcs_b200.so is older than its sources: run `python -m vcs_h264_b200.build`",
                          RuntimeWarning, stacklevel=2)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = header/library mismatch: be loud
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One vcs_ctx: a device, a stream, grow-only scratch.  Not thread-safe."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        rc = self.lib.vcs_create(device, C.byref(h))
        if rc != OK:
            raise VcsError(rc, "vcs_create failed: no usable CUDA device (this package has no "
                               "CPU path)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.vcs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != OK:
            raise VcsError(rc, self.lib.vcs_last_error(self.h).decode())

    def call(self, name, *args):
        self.check(getattr(self.lib, name)(self.h, *args))

    # -- plumbing -------------------------------------------------------------------------
    def set_stream(self, stream_ptr):
        self.call("vcs_set_stream", stream_ptr)

    def use_own_stream(self):
        self.call("vcs_use_own_stream")

    def synchronize(self):
        self.call("vcs_synchronize")

    def device_info(self):
        sm, ma, mi, sh = _i(), _i(), _i(), _sz()
        self.call("vcs_device_info", C.byref(sm), C.byref(ma), C.byref(mi), C.byref(sh))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), smem_optin=sh.value)

    def launch_count(self):
        return int(self.lib.vcs_launch_count(self.h))

    def set_q(self, Q):
        Q = np.ascontiguousarray(np.asarray(Q, np.float64).reshape(3, 64))
        self.call("vcs_set_q", Q.ctypes.data)

    def microbench(self, which, iters=2000):
        r, m = _d(), _d()
        self.call("vcs_microbench", which, iters, C.byref(r), C.byref(m))
        return r.value, m.value

    def enable_kernel_timing(self, on=True):
        self.call("vcs_enable_kernel_timing", int(on))

    def kernel_times(self):
        a, b, n = _d(), _d(), _i()
        self.call("vcs_kernel_times", C.byref(a), C.byref(b), C.byref(n))
        return a.value, b.value, n.value


def me_reference_params(H, W, bs) -> MeParams:
    p = MeParams()
    if load().vcs_me_reference_params(H, W, bs, C.byref(p)) != OK:
        raise ValueError(f"invalid block size {bs}")
    return p


def me_fullsearch_params(H, W, bs, R, metric=METRIC_WRAP8, static_thr=2000) -> MeParams:
    p = MeParams()
    if load().vcs_me_fullsearch_params(H, W, bs, R, metric, static_thr, C.byref(p)) != OK:
        raise ValueError("invalid full-search parameters")
    return p


def q_tables(qf=50.0):
    Q = np.empty((3, 8, 8), np.float64)
    if load().vcs_q_tables(float(qf), Q.ctypes.data) != OK:
        raise ValueError("Invalid quality setting, must be between 1 and 99.")
    return Q


def dct_matrix():
    m = np.empty((8, 8), np.float64)
    load().vcs_dct_matrix(m.ctypes.data)
    return m


def num_blocks(H, W, bs):
    return load().vcs_num_blocks(H, W, bs)


def num_p_frames(T, gop_len):
    return load().vcs_num_p_frames(T, gop_len)
