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


def test_private_dct_helpers_match_the_oracle(orc):
    """DCTCompressor._dct2 / _idct2 (DCTcompressor.py:111-121) on the CUDA path, bit for bit."""
    import vcs_h264_b200 as v
    dc = v.DCTCompressor(8)
    rng = np.random.default_rng(5)
    for _ in range(8):
        x = rng.integers(-128, 128, (8, 8)).astype(np.float64)
        d = dc._dct2(x)
        assert np.array_equal(d, orc.dct2(x))
        assert np.array_equal(dc._idct2(d), orc.idct2(d))
    assert np.array_equal(dc._dctMatrix(), orc.dct_matrix())


def test_out_of_frame_vectors_are_refused_not_followed():
    """*_dev paths never follow a motion vector out of the frame (corrupt or foreign input): the prediction is zero
    there and the context reports it at the next synchronisation."""
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W, bs, gop = 4, 64, 96, 16, 4
    clip = torch.from_numpy(synth.clip(T, H, W, seed=9, margin=48)).cuda()
    ce = v.ClipEncoder([H, W], block_size=bs, search="full", search_range=8, gop_len=gop, coef_mode=v.COEF_I16_RINT)
    out = ce.alloc_device_outputs(T)
    ce.encode_device(clip, out)
    torch.cuda.synchronize()
    cd = v.ClipDecoder([H, W], block_size=bs, gop_len=gop, coef_mode=v.COEF_I16_RINT)
    rec = torch.empty_like(out["recon"])
    cd.decode_device(clip[::gop].contiguous(), out["mv"], out["coef"], rec, T)
    cd.ctx.synchronize()
    assert torch.equal(rec, out["recon"])
    bad = out["mv"].clone()
    bad[1, 7, 0] = 30000                                   # far outside the frame
    bad[2, 0, 1] = -5                                      # above the top edge
    cd.decode_device(clip[::gop].contiguous(), bad, out["coef"], rec, T)
    with pytest.raises(v.VcsError, match="outside the frame"):
        cd.ctx.synchronize()
    cd.ctx.synchronize()                                   # reported once, then clear
    # the good macroblocks are unaffected
    assert torch.equal(rec[0], out["recon"][0])


def test_front_ends_do_not_share_quantiser_state():
    """Two encoders with different QF on one device keep their own tables (each owns a context)."""
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W = 2, 64, 96
    clip = torch.from_numpy(synth.clip(T, H, W, seed=3, margin=48)).cuda()
    a = v.ClipEncoder([H, W], block_size=16, search_range=8, gop_len=2, qf=50.0, coef_mode=v.COEF_I16_RINT)
    oa = a.alloc_device_outputs(T)
    a.encode_device(clip, oa)
    torch.cuda.synchronize()
    want = oa["coef"].clone()
    b = v.ClipEncoder([H, W], block_size=16, search_range=8, gop_len=2, qf=90.0, coef_mode=v.COEF_I16_RINT)
    ob = b.alloc_device_outputs(T)
    b.encode_device(clip, ob)
    v.DCTCompressor(8).compress(np.zeros((8, 8, 3), np.uint8))      # the shared drop-in context, QF 50 tables
    a.encode_device(clip, oa)
    torch.cuda.synchronize()
    assert torch.equal(oa["coef"], want) and not torch.equal(ob["coef"], want)


def test_two_devices_in_one_process():
    """Contexts on two devices interleave their calls; each entry point selects its own device and puts the
    caller's back."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W = 3, 64, 96
    clip = synth.clip(T, H, W, seed=4, margin=48)
    enc = [v.ClipEncoder([H, W], block_size=16, search_range=8, gop_len=3, coef_mode=v.COEF_I16_RINT, device=d) for d in (0, 1)]
    torch.cuda.set_device(0)
    outs = [e.encode_host(clip, want_coef=True, want_recon=True) for e in (enc[1], enc[0], enc[1])]
    assert torch.cuda.current_device() == 0
    for o in outs[1:]:
        for k in ("mv", "coef", "recon"):
            assert np.array_equal(np.asarray(o[k]), np.asarray(outs[0][k])), k
