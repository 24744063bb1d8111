"""4:2:0 chroma subsampling (ChromaSubsampling/chroma.py).

CPU: the oracle against vectors made by the unmodified reference script (tests/golden/make_golden_chroma.py).
GPU: chroma.cuh through the C ABI against the oracle and the same vectors, bit-exact."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["corgi", "odd", "sat", "tiny"]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "golden_chroma.npz"))


@pytest.fixture(scope="module")
def meta():
    with open(os.path.join(HERE, "golden", "golden_chroma_meta.json")) as f:
        return json.load(f)


def test_golden_was_pinned_on_the_reference_script(meta):
    assert meta["oracle_mismatch_full_image"] == {"Y": 0, "cr": 0, "cb": 0, "final": 0}
    assert meta["boxfilter_sums_seen"] == 1021 and meta["boxfilter_is_ceil_quarter"]
    assert meta["corgi_crop_equals_script"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(gold, name):
    Y, cr, cb = orc.chroma420(gold[name + "_img"])
    assert np.array_equal(Y, gold[name + "_Y"])
    assert np.array_equal(cr, gold[name + "_cr"])
    assert np.array_equal(cb, gold[name + "_cb"])
    assert np.array_equal(orc.chroma420_to_bgr(Y, cr, cb), gold[name + "_final"])


def test_sample_geometry_is_the_box_filters_not_the_aligned_block():
    # a single bright pixel at (1, 1) lands in samples (0,0) (by reflection), (0,1), (1,0) and (1,1): windows are
    # rows {2i-1, 2i} x cols {2j-1, 2j}, not the aligned 2x2 block
    img = np.zeros((4, 4, 3), np.uint8)
    img[1, 1] = 255
    _, cr, cb = orc.chroma420(img)
    assert cr.shape == (2, 2) and cb.shape == (2, 2)
    assert (cr != 128).sum() + (cb != 128).sum() == 0   # white pixel is colourless
    img[1, 1] = (0, 0, 255)                              # red: Cr far above 128
    _, cr, _ = orc.chroma420(img)
    assert (cr > 128).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(gold, name):
    from vcs_h264_b200 import chroma
    (Y, cr, cb), final = chroma.subsample420(gold[name + "_img"], with_reconstruction=True)
    assert np.array_equal(Y, gold[name + "_Y"])
    assert np.array_equal(cr, gold[name + "_cr"])
    assert np.array_equal(cb, gold[name + "_cb"])
    assert np.array_equal(final, gold[name + "_final"])
    assert np.array_equal(chroma.reconstruct([Y, cr, cb]), gold[name + "_final"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (5, 1), (2, 2), (3, 3), (64, 64), (129, 257), (720, 1280), (2160, 3840)])
def test_cuda_matches_oracle_random(shape):
    from vcs_h264_b200 import chroma
    rng = np.random.default_rng(shape[0] * 10007 + shape[1])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    (Y, cr, cb), final = chroma.subsample420(img, with_reconstruction=True)
    oY, ocr, ocb = orc.chroma420(img)
    assert np.array_equal(Y, oY) and np.array_equal(cr, ocr) and np.array_equal(cb, ocb)
    assert np.array_equal(final, orc.chroma420_to_bgr(oY, ocr, ocb))


@pytest.mark.gpu
def test_cuda_rejects_bad_input():
    from vcs_h264_b200 import chroma
    with pytest.raises(ValueError):
        chroma.subsample420(np.zeros((4, 4), np.uint8))
    with pytest.raises(ValueError):
        chroma.reconstruct([np.zeros((4, 4), np.uint8), np.zeros((1, 2), np.uint8), np.zeros((2, 2), np.uint8)])
