"""Clip-level (batched) front-end of the hot path.

The reference encodes one frame per Python call (main.py:34-41).  At kfps rates the per-frame
list marshalling would dominate, so this module exposes the same pipeline -- Encoder's GOP rule
(encoder.py:25,51-52), MotionProcessor search, residual, DCTCompressor quantiser and the decoder's
reconstruction -- over a whole clip in one C-ABI call, with array outputs indexed by P-frame
ordinal.  torch is used only for device memory and streams.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .runtime import acquire_context, release_context

COEF_DTYPES = {_capi.COEF_F64: np.float64, _capi.COEF_F64_RINT: np.float64,
               _capi.COEF_I16_RINT: np.int16, _capi.COEF_I8_RINT: np.int8}


class ClipEncoder:
    """ME (+ static test) -> MC -> residual -> 8x8 DCT -> quantise [-> dequantise -> IDCT ->
    reconstruct] for every P-frame of a clip.

    search: "reference" = MotionProcessor's own window/step (motion.py:123-140);
            "full"      = symmetric +/-search_range, step 1 (BASELINE.json configs 2/3/5).
    """

    def __init__(self, shape, block_size=16, search="full", search_range=16, gop_len=4, qf=50.0,
                 metric=_capi.METRIC_WRAP8, static_thr=2000, coef_mode=_capi.COEF_I16_RINT,
                 kernel=_capi.ME_AUTO, device=0, dct_precision=64):
        H, W = int(shape[0]), int(shape[1])
        if search == "reference":
            self.params = _capi.me_reference_params(H, W, block_size)
            self.params.metric = metric
            self.params.static_thr = static_thr
        elif search == "full":
            self.params = _capi.me_fullsearch_params(H, W, block_size, search_range, metric, static_thr)
        else:
            raise ValueError("search must be 'reference' or 'full'")
        self.params.kernel = kernel
        if gop_len < 2:
            raise ValueError("gop_len must be >= 2 (pattern of one I-frame + P-frames)")
        self.H, self.W, self.bs, self.gop_len = H, W, block_size, gop_len
        self.N = _capi.num_blocks(H, W, block_size)
        self.coef_mode = coef_mode
        self.device = device
        # own context: its Q tables, stream and scratch are not shared with other front-ends on the device
        self.ctx = acquire_context(device)
        self.Q = _capi.q_tables(qf)
        self.ctx.set_q(self.Q)
        # 64: float64 transforms, bit-identical with the reference (default).  32: the fp32 tier -- close, not equal;
        # compare with flip_counters().  Set on every construction: contexts are pooled.
        self.dct_precision = int(dct_precision)
        self.ctx.call("vcs_set_dct_precision", self.dct_precision)

    def close(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        release_context(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_p_frames(self, T):
        return _capi.num_p_frames(T, self.gop_len)

    def p_frame_indices(self, T):
        return [t for t in range(T) if t % self.gop_len != 0]

    # -- host buffers in, host buffers out (bench.py e2e; copies inside) -----------------------
    def alloc_host_outputs(self, T, want_coef=True, want_recon=False, pinned=True):
        import torch
        nP = self.num_p_frames(T)

        def buf(shape, dtype):
            t = torch.empty(shape, dtype=dtype)
            return t.pin_memory() if pinned and torch.cuda.is_available() else t
        out = dict(mv=buf((nP, self.N, 2), torch.int16), cost=buf((nP, self.N), torch.int32),
                   flags=buf((nP, self.N), torch.uint8))
        if want_coef:
            dt = {_capi.COEF_I16_RINT: torch.int16, _capi.COEF_I8_RINT: torch.int8}.get(self.coef_mode, torch.float64)
            out["coef"] = buf((nP, 3, self.H, self.W), dt)
        if want_recon:
            out["recon"] = buf((nP, self.H, self.W, 3), torch.uint8)
        return out

    def encode_host(self, frames, out=None, want_coef=True, want_recon=False):
        """frames: uint8 [T,H,W,3] numpy array or (pinned) CPU torch tensor."""
        T = int(frames.shape[0])
        if tuple(frames.shape[1:]) != (self.H, self.W, 3):
            raise ValueError(f"frames must be [T,{self.H},{self.W},3] uint8")
        if out is None:
            out = self.alloc_host_outputs(T, want_coef, want_recon, pinned=False)
        self.ctx.call("vcs_encode_clip_host", self.params, _capi.ptr(frames), T, self.gop_len,
                      self.coef_mode, _capi.ptr(out["mv"]), _capi.ptr(out.get("cost")),
                      _capi.ptr(out.get("flags")), _capi.ptr(out.get("coef")),
                      _capi.ptr(out.get("recon")))
        return out

    # -- host buffers in, PACKED coefficients out (bitmap + 4-bit codes + escapes: about 40 % of the dense int8 bytes) ---
    def alloc_host_packed(self, T, want_recon=False, pinned=True, escape_fraction=0.25):
        """Output buffers of encode_host_packed.  The nibble stream gets its worst-case size; the escape stream
        `escape_fraction` of its worst case (every index outside [-8, 7]); 1.0 can never overflow."""
        import torch
        nP = self.num_p_frames(T)

        def buf(shape, dtype):
            t = torch.empty(shape, dtype=dtype)
            return t.pin_memory() if pinned and torch.cuda.is_available() else t
        ncoef, nblk = nP * 3 * self.H * self.W, nP * 3 * (self.H // 8) * (self.W // 8)
        out = dict(mv=buf((nP, self.N, 2), torch.int16), flags=buf((nP, self.N), torch.uint8),
                   bitmap=buf((nP, 3, self.H // 8, self.W // 8), torch.int64),      # uint64 bit patterns
                   row_count=buf((nP, 3, self.H // 8, 2), torch.int32),
                   nibbles=buf((ncoef // 2 + nblk,), torch.uint8),
                   escapes=buf((max(64, int(ncoef * escape_fraction)),), torch.int8))
        if want_recon:
            out["recon"] = buf((nP, self.H, self.W, 3), torch.uint8)
        return out

    def encode_host_packed(self, frames, out=None, want_recon=False):
        """Like encode_host with int8 indices, but the indices come back packed (include/vcs_b200.h:
        vcs_encode_clip_host_packed).  out["lengths"] = (bytes of the nibble stream, number of escapes)."""
        import ctypes as C
        T = int(frames.shape[0])
        if tuple(frames.shape[1:]) != (self.H, self.W, 3):
            raise ValueError(f"frames must be [T,{self.H},{self.W},3] uint8")
        if out is None:
            out = self.alloc_host_packed(T, want_recon, pinned=False, escape_fraction=1.0)
        n = (C.c_uint64 * 2)(0, 0)

        def size(a):
            return a.numel() if hasattr(a, "numel") else a.size
        self.ctx.call("vcs_encode_clip_host_packed", self.params, _capi.ptr(frames), T, self.gop_len,
                      _capi.ptr(out["mv"]), _capi.ptr(out.get("cost")), _capi.ptr(out.get("flags")),
                      _capi.ptr(out["bitmap"]), _capi.ptr(out["row_count"]), _capi.ptr(out["nibbles"]), size(out["nibbles"]),
                      _capi.ptr(out["escapes"]), size(out["escapes"]), n, _capi.ptr(out.get("recon")))
        out["lengths"] = (int(n[0]), int(n[1]))
        return out

    @staticmethod
    def packed_bytes(out):
        """Bytes of a packed result that carry information (what travels: vectors, flags, bitmaps, counts, streams)."""
        def nbytes(a):
            return a.numel() * a.element_size() if hasattr(a, "numel") else a.nbytes
        return sum(nbytes(out[k]) for k in ("mv", "flags", "bitmap", "row_count") if out.get(k) is not None) + sum(out["lengths"])

    # -- device resident (bench.py value) --------------------------------------------------------
    def alloc_device_outputs(self, T, want_coef=True, want_recon=True):
        import torch
        nP = self.num_p_frames(T)
        dev = torch.device("cuda", self.device)
        out = dict(mv=torch.empty((nP, self.N, 2), dtype=torch.int16, device=dev),
                   cost=torch.empty((nP, self.N), dtype=torch.int32, device=dev),
                   flags=torch.empty((nP, self.N), dtype=torch.uint8, device=dev))
        if want_coef:
            dt = {_capi.COEF_I16_RINT: torch.int16, _capi.COEF_I8_RINT: torch.int8}.get(self.coef_mode, torch.float64)
            out["coef"] = torch.empty((nP, 3, self.H, self.W), dtype=dt, device=dev)
        if want_recon:
            out["recon"] = torch.empty((nP, self.H, self.W, 3), dtype=torch.uint8, device=dev)
        return out

    def encode_device(self, frames_dev, out, stream=None):
        """frames_dev: uint8 CUDA tensor [T,H,W,3]; kernels are enqueued on `stream`
        (default: torch's current stream) and not synchronised."""
        import torch
        T = int(frames_dev.shape[0])
        s = stream if stream is not None else torch.cuda.current_stream(frames_dev.device)
        self.ctx.set_stream(s.cuda_stream)        # borrowed for this call only: the handle may not outlive it
        try:
            self.ctx.call("vcs_encode_clip_dev", self.params, _capi.ptr(frames_dev), T, self.gop_len,
                          self.coef_mode, _capi.ptr(out["mv"]), _capi.ptr(out.get("cost")),
                          _capi.ptr(out.get("flags")), _capi.ptr(out.get("coef")),
                          _capi.ptr(out.get("recon")))
        finally:
            self.ctx.use_own_stream()
        return out

    def me_device(self, frames_dev, out, stream=None):
        """Search only (BASELINE config 5's ME-only sweep)."""
        import torch
        T = int(frames_dev.shape[0])
        s = stream if stream is not None else torch.cuda.current_stream(frames_dev.device)
        self.ctx.set_stream(s.cuda_stream)
        try:
            self.ctx.call("vcs_me_search_clip_dev", self.params, _capi.ptr(frames_dev), T, self.gop_len,
                          _capi.ptr(out["mv"]), _capi.ptr(out.get("cost")), _capi.ptr(out.get("flags")))
        finally:
            self.ctx.use_own_stream()
        return out


class ClipDecoder:
    """Decoder.reconstruct_video's per-frame arithmetic (decoder.py:23-69) for a whole clip: motion
    compensation from the ORIGINAL I-frames, dequantise, IDCT, truncating store, YCrCb->BGR, wrap add."""

    def __init__(self, shape, block_size=16, gop_len=4, qf=50.0, coef_mode=_capi.COEF_I16_RINT, device=0, dct_precision=64):
        self.H, self.W, self.bs, self.gop_len = int(shape[0]), int(shape[1]), block_size, gop_len
        self.coef_mode, self.device = coef_mode, device
        self.ctx = acquire_context(device)         # own context, like ClipEncoder
        self.Q = _capi.q_tables(qf)
        self.ctx.set_q(self.Q)
        self.dct_precision = int(dct_precision)
        self.ctx.call("vcs_set_dct_precision", self.dct_precision)

    def close(self):
        ctx, self.ctx = getattr(self, "ctx", None), None
        release_context(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_host(self, ref_frames, mv, coef, T):
        """ref_frames uint8 [nG,H,W,3] (the I-frames), mv int16 [nP,N,2], coef [nP,3,H,W] -> uint8 [nP,H,W,3]."""
        nP = _capi.num_p_frames(T, self.gop_len)
        mv = np.ascontiguousarray(np.asarray(mv), np.int16)
        coef = np.ascontiguousarray(np.asarray(coef), COEF_DTYPES[self.coef_mode])
        ref_frames = np.ascontiguousarray(np.asarray(ref_frames), np.uint8)
        out = np.empty((nP, self.H, self.W, 3), np.uint8)
        self.ctx.call("vcs_decode_clip_host", self.H, self.W, self.bs, ref_frames.ctypes.data, T, self.gop_len,
                      mv.ctypes.data, self.coef_mode, coef.ctypes.data, out.ctypes.data)
        return out

    def decode_host_packed(self, ref_frames, mv, bitmap, row_count, nibbles, escapes, lengths, T):
        """decode_host from the packed coefficient form (ClipEncoder.encode_host_packed / container version 2)."""
        nP = _capi.num_p_frames(T, self.gop_len)
        mv = np.ascontiguousarray(np.asarray(mv), np.int16)
        ref_frames = np.ascontiguousarray(np.asarray(ref_frames), np.uint8)
        bitmap = np.ascontiguousarray(np.asarray(bitmap)).view(np.uint64)
        row_count = np.ascontiguousarray(np.asarray(row_count)).view(np.uint32)
        nibbles = np.ascontiguousarray(np.asarray(nibbles).reshape(-1)[:lengths[0]]).view(np.uint8)
        escapes = np.ascontiguousarray(np.asarray(escapes).reshape(-1)[:lengths[1]]).view(np.int8)
        out = np.empty((nP, self.H, self.W, 3), np.uint8)
        self.ctx.call("vcs_decode_clip_host_packed", self.H, self.W, self.bs, ref_frames.ctypes.data, T, self.gop_len,
                      mv.ctypes.data, bitmap.ctypes.data, row_count.ctypes.data, nibbles.ctypes.data, int(lengths[0]),
                      escapes.ctypes.data, int(lengths[1]), out.ctypes.data)
        return out

    def decode_device(self, ref_frames_dev, mv_dev, coef_dev, recon_dev, T, stream=None):
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(recon_dev.device)
        self.ctx.set_stream(s.cuda_stream)
        try:
            self.ctx.call("vcs_decode_clip_dev", self.H, self.W, self.bs, _capi.ptr(ref_frames_dev), T, self.gop_len,
                          _capi.ptr(mv_dev), self.coef_mode, _capi.ptr(coef_dev), _capi.ptr(recon_dev))
        finally:
            self.ctx.use_own_stream()
        return recon_dev


def flip_counters(ctx, coef_mode, coef_a, coef_b, recon_a=None, recon_b=None):
    """How two device results differ (the fp32 tier against the exact one): rounded indices that crossed a rounding
    boundary, pixels that crossed a truncation boundary, and the PSNR of one reconstruction against the other."""
    import ctypes as C
    import math
    out = (C.c_ulonglong * 3)()
    npx = recon_a.numel() if recon_a is not None else 0
    ctx.call("vcs_flip_counters_dev", coef_mode, _capi.ptr(coef_a), _capi.ptr(coef_b), coef_a.numel(),
             _capi.ptr(recon_a) if npx else None, _capi.ptr(recon_b) if npx else None, npx, out)
    mse = out[2] / npx if npx else 0.0
    return {"indices": coef_a.numel(), "index_flips": int(out[0]), "pixels": npx, "pixel_flips": int(out[1]),
            "psnr_db": (10.0 * math.log10(255.0 ** 2 / mse) if mse > 0 else float("inf")) if npx else None}


def sparsity_device(ctx, coef_dev, coef_mode):
    """1 - nnz/size of a device coefficient tensor (the print of DCTCompression/dct.py:188-191)."""
    import ctypes as C
    cnt = C.c_ulonglong(0)
    n = coef_dev.numel()
    ctx.call("vcs_count_nonzero_dev", coef_mode, _capi.ptr(coef_dev), n, C.byref(cnt))
    return 1.0 - cnt.value / float(n)
