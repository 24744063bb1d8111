"""Out-of-bounds write guards (compute-sanitizer is not available on this pool): every device output of the clip
encoder, the decoder, the intra and the chroma kernels is carved out of one arena with 0xA5-filled gaps on both sides;
after the kernels ran the gaps must be untouched, and the results must equal a run into ordinary tensors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GUARD = 4096


class Arena:
    def __init__(self, torch, nbytes):
        self.torch = torch
        self.buf = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device="cuda")
        self.off = GUARD
        self.spans = []

    def take(self, shape, dtype):
        torch = self.torch
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        start = (self.off + 255) // 256 * 256
        view = self.buf[start:start + n].view(dtype).view(*shape)
        self.spans.append((start, start + n))
        self.off = start + n + GUARD
        assert self.off <= self.buf.numel()
        return view

    def check(self):
        mask = self.torch.ones(self.buf.numel(), dtype=self.torch.bool, device="cuda")
        for a, b in self.spans:
            mask[a:b] = False
        assert bool((self.buf[mask] == 0xA5).all()), "a kernel wrote outside its output buffer"


@pytest.mark.parametrize("geom", [(1080 // 4 + 2, 1920 // 4, 16, 16, 3), (72, 104, 8, 8, 1), (64, 100, 4, 4, 2), (80, 112, 16, 32, 0)])
def test_clip_encoder_and_decoder_stay_inside_their_buffers(geom):
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    H, W, bs, R, cm = geom
    H, W = H // 8 * 8, W // 8 * 8                       # the DCT stage needs multiples of 8
    T, gop = 7, 3
    clip = torch.from_numpy(synth.clip(T, H, W, seed=H * W, margin=64)).cuda()
    ce = v.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=gop, coef_mode=cm)
    ref = ce.alloc_device_outputs(T)
    ce.encode_device(clip, ref)
    torch.cuda.synchronize()
    arena = Arena(torch, sum(t.numel() * t.element_size() for t in ref.values()) + 16 * GUARD)
    out = {k: arena.take(tuple(t.shape), t.dtype) for k, t in ref.items()}
    ce.encode_device(clip, out)
    torch.cuda.synchronize()
    arena.check()
    for k in ref:
        assert torch.equal(out[k], ref[k]), k
    # decoder side into a guarded reconstruction buffer
    cd = v.ClipDecoder([H, W], block_size=bs, gop_len=gop, coef_mode=cm)
    arena2 = Arena(torch, ref["recon"].numel() + 4 * GUARD)
    rec = arena2.take(tuple(ref["recon"].shape), torch.uint8)
    cd.decode_device(clip[::gop].contiguous(), ref["mv"], ref["coef"], rec, T)
    torch.cuda.synchronize()
    arena2.check()
    assert torch.equal(rec, ref["recon"])


@pytest.mark.parametrize("shape", [(48, 64), (1, 1), (37, 53), (2, 3)])
def test_chroma_and_intra_stay_inside_their_buffers(shape):
    import torch
    from vcs_h264_b200 import _capi, runtime
    H, W = shape
    rng = np.random.default_rng(H * 131 + W)
    img = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).cuda()
    ctx = runtime.get_context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        h2, w2 = (H + 1) // 2, (W + 1) // 2
        arena = Arena(torch, H * W * 4 + 2 * h2 * w2 + 16 * GUARD + H * W * 16 * 4)
        Y, cr, cb = arena.take((H, W), torch.uint8), arena.take((h2, w2), torch.uint8), arena.take((h2, w2), torch.uint8)
        back = arena.take((H, W, 3), torch.uint8)
        ctx.call("vcs_chroma420_dev", H, W, _capi.ptr(img), _capi.ptr(Y), _capi.ptr(cr), _capi.ptr(cb))
        ctx.call("vcs_chroma420_to_bgr_dev", H, W, _capi.ptr(Y), _capi.ptr(cr), _capi.ptr(cb), _capi.ptr(back))
        if H % 16 == 0 and W % 16 == 0:
            plane = img[..., 0].contiguous()
            plane2 = img[..., 1].contiguous()
            o = [arena.take((H, W), torch.int32) for _ in range(4)]
            for name, m, args in (("vcs_intra_luma4x4_dev", 4, (plane, o[0], o[1])), ("vcs_intra_luma16x16_dev", 16, (plane, o[0], o[1])),
                                  ("vcs_intra_chroma8x8_dev", 8, (plane, plane2, o[0], o[1], o[2], o[3]))):
                modes = arena.take((H // m, W // m), torch.uint8)
                ctx.call(name, H, W, *[_capi.ptr(a) for a in args], _capi.ptr(modes))
        torch.cuda.synchronize()
        arena.check()
    finally:
        ctx.use_own_stream()


def test_misaligned_device_buffers_are_refused():
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import _capi, runtime
    H, W = 16, 32
    ctx = runtime.get_context(0)
    img = torch.zeros(H * W * 3 + 8, dtype=torch.uint8, device="cuda")
    coef = torch.zeros(3 * H * W * 2 + 64, dtype=torch.uint8, device="cuda")
    ctx.call("vcs_compress_dev", H, W, img.data_ptr(), _capi.COEF_I16_RINT, coef.data_ptr())          # aligned: fine
    with pytest.raises(v.VcsError):
        ctx.call("vcs_compress_dev", H, W, img.data_ptr() + 1, _capi.COEF_I16_RINT, coef.data_ptr())
    with pytest.raises(v.VcsError):
        ctx.call("vcs_compress_dev", H, W, img.data_ptr(), _capi.COEF_I16_RINT, coef.data_ptr() + 8)
    torch.cuda.synchronize()
