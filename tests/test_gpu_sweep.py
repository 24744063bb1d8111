"""BASELINE config 5 (search-range sweep +/-4 .. +/-64 at 720p / 1080p / 4K, ME only): at every point of the sweep the
tiled kernel's motion vectors and costs of the top four macroblock rows equal the CPU oracle's, for both costs.
(tools/sweep.py measures the same points; the parity check lives here.)"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RES = {"720p": (720, 1280), "1080p": (1080, 1920), "4K": (2160, 3840)}


@pytest.fixture(scope="module")
def clips():
    from vcs_h264_b200 import synth
    return {name: synth.clip(2, H, W, seed=7, margin=96) for name, (H, W) in RES.items()}


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("R", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("res", list(RES))
def test_sweep_point_strip_is_bit_exact(clips, res, R, metric):
    import torch
    import vcs_h264_b200 as v
    H, W = RES[res]
    clip_np = clips[res]
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=R, gop_len=4, metric=metric,
                       static_thr=-1, kernel=v.ME_TILED)
    out = ce.alloc_device_outputs(2, want_coef=False, want_recon=False)
    ce.me_device(torch.from_numpy(clip_np).cuda(), out)
    torch.cuda.synchronize()
    strip = 64 + R + 16                                     # the top 4 MB rows only see reference rows < 64 + R + 16
    omv, ocost, _ = orc.me(clip_np[1][:strip], clip_np[0][:strip], 16, metric=metric, static_thr=-1,
                           **orc.symmetric_search_params(R))
    n = (64 // 16) * (W // 16)
    assert np.array_equal(out["mv"][0].cpu().numpy().astype(np.int32)[:n], omv[:n])
    assert np.array_equal(out["cost"][0].cpu().numpy().view(np.uint32)[:n], ocost[:n])
