#!/usr/bin/env python
"""Pins for BASELINE.json configs[0]: the UNMODIFIED reference on its own default input.

Runs InterframeCompression/main.py's encode (main.py:13-16,29-41: videos/traffic_cut.mp4, BLOCK_SIZE 8,
["I","P","P","P"], motion + residual DCT) and its decoder's per-frame arithmetic (decoder.py:52-69) through the
reference's own classes, imported as they lie under /root/reference, and records

    tests/golden/traffic_cut.mp4          the input clip (a reference DATA fixture, copied byte for byte so that the
                                          GPU box, which has no /root/reference, can decode the same file)
    tests/golden/golden_traffic_meta.json sha256 pins: decoded input frames, all 114x3600 motion vectors, static
                                          counts, and for P-frames 1 and 35 the float64 coefficient planes and the
                                          decoder's final frames

Run in the build container only:  python tests/golden/make_golden_traffic.py
"""
import contextlib
import hashlib
import io
import json
import os
import shutil
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

sys.modules.setdefault("matplotlib", MagicMock())
sys.modules.setdefault("matplotlib.pyplot", MagicMock())
sys.path.insert(0, os.path.join(REF, "InterframeCompression"))
import cv2  # noqa: E402
from encoder import Encoder  # noqa: E402
from decoder import Decoder  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    src = os.path.join(REF, "videos", "traffic_cut.mp4")
    dst = os.path.join(HERE, "traffic_cut.mp4")
    shutil.copyfile(src, dst)
    cap = cv2.VideoCapture(dst)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    cap.release()
    H, W = frames[0].shape[:2]
    meta = dict(file="traffic_cut.mp4", file_sha256=hashlib.sha256(open(dst, "rb").read()).hexdigest(),
                frames=len(frames), shape=[H, W], frames_sha16=sha(np.stack(frames)),
                cv2=cv2.__version__, numpy=np.__version__)
    enc = Encoder(["I", "P", "P", "P"], [H, W], 8, True)        # main.py:29-30
    with contextlib.redirect_stdout(io.StringIO()):
        for n, f in enumerate(frames):                           # main.py:34-41
            enc.encode_frame(f, n)
    P = [f for f in enc.encoded_frames if f.t == "P"]
    mv = np.asarray([f.mv for f in P], np.int32)
    meta["n_p"] = len(P)
    meta["mv_sha16"] = sha(mv)
    statics = [int(((m[:, 0] == 0) & (m[:, 1] == 0)).sum()) for m in mv]
    meta["static_minmaxmean"] = [min(statics), max(statics), float(np.mean(statics))]
    meta["coords_sha16"] = sha(np.asarray(P[0].c, np.int32))
    dec = Decoder(enc.encoded_frames, 25, [H, W], enc.ref_frames, 8, True)
    psnr = []
    for n in (1, 35):
        fr = enc.encoded_frames[n]
        assert fr.t == "P" and fr.i == n
        meta[f"frame{n}_planes_sha16"] = sha(np.stack(fr.r))
        with contextlib.redirect_stdout(io.StringIO()):
            final = dec._reconstruct_P_frame(fr, True)
        meta[f"frame{n}_final_sha16"] = sha(final)
        err = final.astype(np.float64) - frames[n].astype(np.float64)
        psnr.append(float(10 * np.log10(255.0 ** 2 / np.mean(err ** 2))))
    meta["psnr_frames_1_35"] = psnr
    # the first 9 P-frames' PSNR
    ps = []
    for fr in P[:9]:
        with contextlib.redirect_stdout(io.StringIO()):
            final = dec._reconstruct_P_frame(fr, True)
        err = final.astype(np.float64) - frames[fr.i].astype(np.float64)
        ps.append(round(float(10 * np.log10(255.0 ** 2 / np.mean(err ** 2))), 2))
    meta["psnr_first9"] = ps
    with open(os.path.join(HERE, "golden_traffic_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
