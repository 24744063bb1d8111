#!/usr/bin/env python
"""Golden vectors for the intra mode decision from the UNMODIFIED reference
(IntraframeCompression/intraframe.py, intramodes.py).  Build container only.

intraframe.py runs `intraframe('../images/happy-corgi.jpg')` at import, so it is imported with
cwd = its directory and matplotlib mocked; the functions are then called on small planes."""
import contextlib
import io
import json
import os
import sys
import warnings
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.modules.setdefault("matplotlib", MagicMock())
sys.modules.setdefault("matplotlib.pyplot", MagicMock())
import cv2  # noqa: E402

cwd = os.getcwd()
os.chdir(os.path.join(REF, "IntraframeCompression"))
sys.path.insert(0, os.getcwd())
with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    import intraframe as ref_intra  # noqa: E402  (executes the corgi run once)
os.chdir(cwd)


def run(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a)


def main():
    out, meta = {}, {}
    img = cv2.imread(os.path.join(REF, "images", "happy-corgi.jpg"))
    img = cv2.resize(img, (16 * (img.shape[1] // 16), 16 * (img.shape[0] // 16)))
    ycc = cv2.cvtColor(img, cv2.COLOR_BGR2YCR_CB)
    Yf, Crf, Cbf = cv2.split(ycc)
    meta["corgi_shape"] = list(Yf.shape)
    # full-image mode histograms (pins of SURVEY 4) + oracle agreement on the whole image
    from oracle import oracle as orc
    r4 = run(ref_intra.luma4x4, Yf)
    r16 = run(ref_intra.luma16x16, Yf)
    rc = run(ref_intra.chroma8x8, Crf, Cbf)
    meta["corgi_hist_luma4x4"] = np.bincount(r4[2].astype(int).ravel(), minlength=9).tolist()
    meta["corgi_hist_luma16x16"] = np.bincount(r16[2].astype(int).ravel(), minlength=3).tolist()
    meta["corgi_hist_chroma8x8"] = np.bincount(rc[4].astype(int).ravel(), minlength=3).tolist()
    o4, o16, oc = orc.luma4x4(Yf), orc.luma16x16(Yf), orc.chroma8x8(Crf, Cbf)
    meta["corgi_full_mismatch"] = {
        "luma4x4": [int((np.asarray(a) != b).sum()) for a, b in zip(r4, o4)],
        "luma16x16": [int((np.asarray(a) != b).sum()) for a, b in zip(r16, o16)],
        "chroma8x8": [int((np.asarray(a) != b).sum()) for a, b in zip(rc, oc)]}
    print(meta)
    # stored cases: crops + adversarial planes
    rng = np.random.default_rng(5)
    cases = {"corgi": (Yf[200:264, 300:380].copy(), Crf[200:264, 300:380].copy(), Cbf[200:264, 300:380].copy()),
             "random": tuple(rng.integers(0, 256, (48, 64), dtype=np.uint8) for _ in range(3)),
             "bright": tuple(rng.integers(200, 256, (32, 48), dtype=np.uint8) for _ in range(3)),   # 3*x wraps
             "flat": tuple(np.full((32, 32), v, np.uint8) for v in (255, 0, 128)),
             "ramp": tuple(((np.add.outer(np.arange(32) * k, np.arange(48) * 5)) % 256).astype(np.uint8) for k in (3, 7, 11))}
    for name, (Y, Cr, Cb) in cases.items():
        out[f"{name}_Y"], out[f"{name}_Cr"], out[f"{name}_Cb"] = Y, Cr, Cb
        a = run(ref_intra.luma4x4, Y)
        out[f"{name}_l4_res"], out[f"{name}_l4_pred"], out[f"{name}_l4_modes"] = [np.asarray(x) for x in a]
        if Y.shape[0] % 16 == 0 and Y.shape[1] % 16 == 0:
            a = run(ref_intra.luma16x16, Y)
            out[f"{name}_l16_res"], out[f"{name}_l16_pred"], out[f"{name}_l16_modes"] = [np.asarray(x) for x in a]
        a = run(ref_intra.chroma8x8, Cr, Cb)
        for k, x in zip(("crres", "crpred", "cbres", "cbpred", "modes"), a):
            out[f"{name}_c8_{k}"] = np.asarray(x)
    meta["cases"] = list(cases)
    np.savez_compressed(os.path.join(HERE, "golden_intra.npz"), **out)
    with open(os.path.join(HERE, "golden_intra_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote golden_intra.npz", os.path.getsize(os.path.join(HERE, "golden_intra.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
