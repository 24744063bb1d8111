"""Intra mode decision: oracle vs the reference's golden vectors (CPU) and the CUDA kernels
vs both (GPU).  All outputs are integers: bit-exact."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gi():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_intra.npz"))


@pytest.fixture(scope="module")
def gi_meta():
    with open(os.path.join(ROOT, "tests", "golden", "golden_intra_meta.json")) as f:
        return json.load(f)


def _check(fns, gi, names):
    luma4, luma16, chroma = fns
    for n in names:
        Y, Cr, Cb = gi[f"{n}_Y"], gi[f"{n}_Cr"], gi[f"{n}_Cb"]
        r = luma4(Y)
        for got, key in zip(r, ("res", "pred", "modes")):
            assert np.array_equal(np.asarray(got).astype(np.float64), gi[f"{n}_l4_{key}"]), (n, "l4", key)
        if f"{n}_l16_res" in gi:
            r = luma16(Y)
            for got, key in zip(r, ("res", "pred", "modes")):
                assert np.array_equal(np.asarray(got).astype(np.float64), gi[f"{n}_l16_{key}"]), (n, "l16", key)
        r = chroma(Cr, Cb)
        for got, key in zip(r, ("crres", "crpred", "cbres", "cbpred", "modes")):
            assert np.array_equal(np.asarray(got).astype(np.float64), gi[f"{n}_c8_{key}"]), (n, "c8", key)


def test_oracle_intra_vs_reference(orc, gi, gi_meta):
    _check((orc.luma4x4, orc.luma16x16, orc.chroma8x8), gi, gi_meta["cases"])


def test_intra_full_image_pins(gi_meta):
    """Recorded in the build container on images/happy-corgi.jpg (736x736): the SURVEY 4 histograms and
    zero mismatches between oracle and reference on every output plane."""
    assert gi_meta["corgi_hist_luma4x4"] == [7986, 11798, 292, 403, 2933, 2491, 3324, 4154, 475]
    assert gi_meta["corgi_hist_luma16x16"] == [668, 946, 502]
    assert gi_meta["corgi_hist_chroma8x8"] == [1, 8441, 22]
    assert all(v == 0 for vs in gi_meta["corgi_full_mismatch"].values() for v in vs)


@pytest.mark.gpu
def test_gpu_intra_vs_reference(gi, gi_meta):
    from vcs_h264_b200 import intraframe
    _check((intraframe.luma4x4, intraframe.luma16x16, intraframe.chroma8x8), gi, gi_meta["cases"])


@pytest.mark.gpu
def test_gpu_intra_4k_vs_oracle(orc):
    """BASELINE config 4 geometry: a synthetic 2160x3840 still, every plane vs the oracle."""
    from vcs_h264_b200 import intraframe, synth
    img = synth.still(2160, 3840, seed=9)
    ycc = orc.bgr2ycrcb(img)
    Y, Cr, Cb = (np.ascontiguousarray(ycc[..., k]) for k in range(3))
    for got, want in zip(intraframe.luma4x4(Y), orc.luma4x4(Y)):
        assert np.array_equal(got, want.astype(np.float64))
    for got, want in zip(intraframe.luma16x16(Y), orc.luma16x16(Y)):
        assert np.array_equal(got, want.astype(np.float64))
    for got, want in zip(intraframe.chroma8x8(Cr, Cb), orc.chroma8x8(Cr, Cb)):
        assert np.array_equal(got, want.astype(np.float64))
