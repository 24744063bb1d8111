"""The fp32 tier of the DCT stage (include/vcs_b200.h: vcs_set_dct_precision, vcs_flip_counters_dev; SURVEY appendix A,
tier T2): same kernel in float, judged by counted flips and PSNR against the exact float64 tier -- never by equality.
The exact tier stays the default and is what every other test pins bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("coef_mode_name", ["COEF_I8_RINT", "COEF_I16_RINT"])
def test_fp32_tier_is_close_and_its_flips_are_counted(coef_mode_name):
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    cm = getattr(v, coef_mode_name)
    T, H, W = 9, 360, 640
    clip = torch.from_numpy(synth.clip(T, H, W, seed=21, margin=48)).cuda()
    outs = {}
    for bits in (64, 32):
        ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=8, gop_len=4, qf=50.0, coef_mode=cm,
                           dct_precision=bits)
        out = ce.alloc_device_outputs(T)
        ce.encode_device(clip, out)
        torch.cuda.synchronize()
        outs[bits] = (ce, out)
    (ce64, a), (ce32, b) = outs[64], outs[32]
    assert torch.equal(a["mv"], b["mv"]) and torch.equal(a["cost"], b["cost"])          # the search is integer work
    fc = v.flip_counters(ce64.ctx, cm, a["coef"], b["coef"], a["recon"], b["recon"])
    # cross-check the device counters with torch
    assert fc["index_flips"] == int((a["coef"] != b["coef"]).sum())
    assert fc["pixel_flips"] == int((a["recon"] != b["recon"]).sum())
    assert fc["indices"] == a["coef"].numel() and fc["pixels"] == a["recon"].numel()
    # close: a handful of rounding-boundary flips (float error ~1e-4 of a quantisation step), each of +-1
    assert fc["index_flips"] < 1e-3 * fc["indices"], fc
    d = (a["coef"].to(torch.int32) - b["coef"].to(torch.int32)).abs()
    assert int(d.max()) <= 1, fc
    # pixels: the reference TRUNCATES the IDCT output (DCTcompressor.py:81,88), and sparse blocks decode to exact integers,
    # so float noise moves many pixels by one level (up to ~50 %); the tier is judged by PSNR, not by count
    assert fc["pixel_flips"] < fc["pixels"], fc
    assert fc["psnr_db"] > 40.0, fc
    # a second exact run is still bit-identical with the first: the switch does not leak between contexts
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=8, gop_len=4, qf=50.0, coef_mode=cm)
    c = ce.alloc_device_outputs(T)
    ce.encode_device(clip, c)
    torch.cuda.synchronize()
    assert torch.equal(c["coef"], a["coef"]) and torch.equal(c["recon"], a["recon"])


def test_fp32_tier_refuses_what_belongs_to_the_exact_tier():
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W = 5, 64, 96
    clip = synth.clip(T, H, W, seed=3, margin=48)
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=8, gop_len=4, qf=50.0, coef_mode=v.COEF_I8_RINT,
                       dct_precision=32)
    with pytest.raises(v.VcsError, match="exact tier"):
        ce.encode_host_packed(clip)
    with pytest.raises(v.VcsError, match="precision"):
        ce.ctx.call("vcs_set_dct_precision", 16)
