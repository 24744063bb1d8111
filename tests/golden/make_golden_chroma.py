#!/usr/bin/env python
"""Golden vectors for the 4:2:0 chroma demo from the UNMODIFIED reference script
ChromaSubsampling/chroma.py.  Build container only (reads /root/reference).

chroma.py is a script: it reads '../images/happy-corgi.jpg', runs at import and writes Output.jpg into the
cwd, so it is executed with runpy from a scratch directory whose parent holds an `images` symlink, with
matplotlib mocked.  Its globals (imgYYC, crSamples, cbSamples, finalImg) are the reference outputs.
A second run feeds it a crafted image (every 2x2 sum 0..1020 on the chroma planes cannot be forced through
cvtColor, so cv2.boxFilter itself is tabulated on a crafted plane) to pin the box filter's rounding."""
import contextlib
import hashlib
import io
import json
import os
import runpy
import sys
import tempfile
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.modules.setdefault("matplotlib", MagicMock())
sys.modules.setdefault("matplotlib.pyplot", MagicMock())
import cv2  # noqa: E402

from oracle import oracle as orc  # noqa: E402


def run_reference_script():
    with tempfile.TemporaryDirectory() as top:
        os.symlink(os.path.join(REF, "images"), os.path.join(top, "images"))
        work = os.path.join(top, "ChromaSubsampling")
        os.mkdir(work)
        cwd = os.getcwd()
        os.chdir(work)
        try:
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                g = runpy.run_path(os.path.join(REF, "ChromaSubsampling", "chroma.py"))
        finally:
            os.chdir(cwd)
    return g


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    g = run_reference_script()
    img, ycc, crS, cbS, final = g["img"], g["imgYYC"], g["crSamples"], g["cbSamples"], g["finalImg"]
    Y, ocr, ocb = orc.chroma420(img)
    ofinal = orc.chroma420_to_bgr(Y, ocr, ocb)
    meta = {"numpy": np.__version__, "opencv": cv2.__version__, "corgi_shape": list(img.shape),
            "corgi_sha": {"img": sha(img), "Y": sha(ycc[:, :, 0]), "cr": sha(crS), "cb": sha(cbS), "final": sha(final)},
            "oracle_mismatch_full_image": {"Y": int((Y != ycc[:, :, 0]).sum()), "cr": int((ocr != crS).sum()),
                                           "cb": int((ocb != cbS).sum()), "final": int((ofinal != final).sum())}}
    # box filter rounding: every 2x2 sum 0..1020 (interior) + the reflected first row/column, odd sizes
    tab = {}
    rng = np.random.default_rng(7)
    for _ in range(40):
        lo = int(rng.integers(0, 256)); hi = int(rng.integers(lo, 256)) + 1
        p = rng.integers(lo, hi, (61, 47), dtype=np.uint8)
        o = cv2.boxFilter(p, ddepth=-1, ksize=(2, 2)).astype(np.int64)
        P = np.pad(p.astype(np.int64), ((1, 0), (1, 0)), mode="reflect")
        s = P[:-1, :-1] + P[:-1, 1:] + P[1:, :-1] + P[1:, 1:]
        for a, b in zip(s.ravel(), o.ravel()):
            assert tab.setdefault(int(a), int(b)) == int(b), "boxFilter is not a function of the 2x2 sum"
    for t in range(1021):                            # every possible sum, deterministically
        q, r = divmod(t, 4)
        p = np.full((4, 4), q, np.uint8)
        for k in range(r):
            p[1 + k // 2, 1 + k % 2] += 1
        v = int(cv2.boxFilter(p, ddepth=-1, ksize=(2, 2))[2, 2])
        assert tab.setdefault(t, v) == v
    meta["boxfilter_sums_seen"] = len(tab)
    meta["boxfilter_is_ceil_quarter"] = all(v == (s + 3) >> 2 for s, v in tab.items())
    out = {}
    # stored cases: a crop of the corgi (top-left corner keeps the reflected border), odd-sized random, extremes
    cases = {"corgi": img[:96, :128], "odd": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
             "sat": np.stack(list(np.meshgrid(np.arange(0, 256, 8), np.arange(0, 256, 8))) + [np.full((32, 32), 255)], -1).astype(np.uint8),
             "tiny": rng.integers(0, 256, (2, 3, 3), dtype=np.uint8)}
    for name, im in cases.items():
        im = np.ascontiguousarray(im)
        yy = cv2.cvtColor(im, cv2.COLOR_BGR2YCR_CB)
        cr = cv2.boxFilter(yy[:, :, 1], ddepth=-1, ksize=(2, 2))[::2, ::2]
        cb = cv2.boxFilter(yy[:, :, 2], ddepth=-1, ksize=(2, 2))[::2, ::2]
        H, W = im.shape[:2]
        fin = np.zeros((H, W, 3), np.uint8)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(H):                      # chroma.py:27-41 verbatim arithmetic
                for j in range(W):
                    Yv = yy[:, :, 0][i, j]
                    Cr = cr[int(i / 2), int(j / 2)]
                    Cb = cb[int(i / 2), int(j / 2)]
                    r = Yv + 1.4022 * (Cr - 128)
                    gg = Yv - 0.34414 * (Cb - 128) - 0.71414 * (Cr - 128)
                    b = Yv + 1.77200 * (Cb - 128)
                    fin[i, j] = [max(0, min(255, b)), max(0, min(255, gg)), max(0, min(255, r))]
        out[name + "_img"], out[name + "_Y"], out[name + "_cr"], out[name + "_cb"], out[name + "_final"] = \
            im, yy[:, :, 0].copy(), cr.copy(), cb.copy(), fin
    # the crop of the reference script's own run must agree with the crop recomputed above wherever the
    # crop's border does not matter (everything except the last row/col of the samples)
    meta["corgi_crop_equals_script"] = bool(np.array_equal(out["corgi_cr"][:47, :63], crS[:47, :63]) and
                                            np.array_equal(out["corgi_final"][:94, :126], final[:94, :126]))
    np.savez_compressed(os.path.join(HERE, "golden_chroma.npz"), **out)
    with open(os.path.join(HERE, "golden_chroma_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
