#!/usr/bin/env python
"""BASELINE.json configs 3 and 5: ME-only search-range sweep (+/-4 .. +/-64 at 720p / 1080p / 4K)
and the 4K +/-32 clip, on one GPU.  Prints one JSON object; bench.py stays the headline line.

    python tools/sweep.py [--frames T] [--metric wrap8|sad]
Timing only: the bit-exactness of every sweep point against the CPU oracle is tests/test_gpu_sweep.py."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def px_ops(H, W, bs, R):
    def n_axis(dim):
        return sum(min(p + R, dim - bs) - max(p - R, 0) + 1 for p in range(0, dim - bs + 1, bs))
    return n_axis(W) * n_axis(H) * 3 * bs * bs


def main():
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--metric", default="sad", choices=["wrap8", "sad"])
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    metric = v.METRIC_SAD if args.metric == "sad" else v.METRIC_WRAP8
    ctx = v.runtime.get_context(0)
    peak = ctx.microbench(0, 4000)[0] * 32 * 4
    rows = []
    for name, (H, W) in (("720p", (720, 1280)), ("1080p", (1080, 1920)), ("4K", (2160, 3840))):
        clip_np = synth.clip(args.frames, H, W, seed=7, margin=96)
        clip = torch.from_numpy(clip_np).cuda()
        for R in (4, 8, 16, 32, 64):
            ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=R, gop_len=4, metric=metric,
                               static_thr=-1, kernel=v.ME_TILED)
            out = ce.alloc_device_outputs(args.frames, want_coef=False, want_recon=False)
            ce.me_device(clip, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                ce.me_device(clip, out)
            e1.record()
            torch.cuda.synchronize()
            nP = ce.num_p_frames(args.frames)
            ms = e0.elapsed_time(e1) / args.iters
            work = px_ops(H, W, 16, R) * nP
            row = {"res": name, "R": R, "p_frames": nP, "ms_per_launch": ms, "us_per_p_frame": 1e3 * ms / nP,
                   "p_frames_per_s": nP / (ms * 1e-3), "Gpxop_per_s": work / (ms * 1e-3) / 1e9,
                   "frac_of_sad_peak": work / (ms * 1e-3) / peak}
            rows.append(row)
            print(json.dumps(row), file=sys.stderr)
    print(json.dumps({"metric": args.metric, "sad_peak_Gpxop_per_s": peak / 1e9, "rows": rows}))


if __name__ == "__main__":
    main()
