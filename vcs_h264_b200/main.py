"""Driver -- the loop of InterframeCompression/main.py:18-57 on the CUDA path.

    python -m vcs_h264_b200.main --video ../videos/traffic_cut.mp4 [--block-size 8] [--no-write]

Reads a video with OpenCV, encodes every frame (I-P-P-P, residual DCT), decodes and -- like the
reference -- writes output.mp4 in the cwd.  `run(frames, ...)` is the same for frames already in memory.
Two encoders are offered: "frame" = the drop-in Encoder/Decoder classes, one call per frame exactly like
the reference; "clip" = ClipEncoder/ClipDecoder, one C-ABI call per clip (same numbers)."""
from __future__ import annotations

import argparse

import numpy as np

VIDEO_INPUT = "../videos/traffic_cut.mp4"      # main.py:9,13
FRAME_RATE = 25
BLOCK_SIZE = 8
ENCODING_PATTERN = ["I", "P", "P", "P"]


def read_video(path):
    import cv2
    cap = cv2.VideoCapture(path)
    frames = []
    while cap.isOpened():
        ret, frame = cap.read()
        if not ret:
            print("Can't receive frame (stream end?). Exiting ...")
            break
        frames.append(frame)
    cap.release()
    return frames


def run(frames, block_size=BLOCK_SIZE, pattern=ENCODING_PATTERN, with_residual=True, with_dct=True, mode="frame"):
    """Encode + decode; returns (encoder-side records, decoded frames)."""
    from . import ClipDecoder, ClipEncoder, Decoder, Encoder, _capi
    H, W = frames[0].shape[:2]
    if mode == "frame":
        enc = Encoder(pattern=pattern, shape=[H, W], block_size=block_size, with_DCT=with_dct and with_residual,
                      dct_block_size=8)
        for n, f in enumerate(frames):
            enc.encode_frame(f, n)
        print("Finished encoding all frames, will decocode and output video.")
        dec = Decoder(encoded_frames=enc.encoded_frames, fps=float(FRAME_RATE), shape=[H, W],
                      ref_frames=enc.ref_frames, block_size=block_size, with_DCT=with_dct and with_residual,
                      dct_block_size=8)
        return enc.encoded_frames, dec.decode_frames(with_residuals=with_residual)
    g = len(pattern)
    clip = np.ascontiguousarray(np.stack(frames))
    ce = ClipEncoder([H, W], block_size=block_size, search="reference", gop_len=g, coef_mode=_capi.COEF_F64)
    out = ce.encode_host(clip, want_coef=True, want_recon=False)
    cd = ClipDecoder([H, W], block_size=block_size, gop_len=g, coef_mode=_capi.COEF_F64)
    rec = cd.decode_host(clip[::g], np.asarray(out["mv"]), np.asarray(out["coef"]), len(frames))
    decoded, p = [], 0
    for t in range(len(frames)):
        if t % g == 0:
            decoded.append(frames[t])
        else:
            decoded.append(rec[p])
            p += 1
    return out, decoded


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--video", default=VIDEO_INPUT)
    ap.add_argument("--block-size", type=int, default=BLOCK_SIZE)
    ap.add_argument("--mode", default="clip", choices=["frame", "clip"])
    ap.add_argument("--no-write", action="store_true")
    a = ap.parse_args(argv)
    frames = read_video(a.video)
    _, decoded = run(frames, block_size=a.block_size, mode=a.mode)
    if not a.no_write:
        import cv2
        H, W = frames[0].shape[:2]
        out = cv2.VideoWriter("output.mp4", cv2.VideoWriter_fourcc(*"X264"), float(FRAME_RATE), (W, H))
        for f in decoded:
            out.write(f)
        out.release()
    print("Releasing everything. Job finished. ")
    print("Finished!")


if __name__ == "__main__":
    main()
