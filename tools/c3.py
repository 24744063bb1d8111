#!/usr/bin/env python
"""BASELINE config 3 geometry on one GPU: synthetic 4K (2160x3840) clip, 16x16 macroblocks, +/-32 step-1 full search
with the reference's wrapped cost + static test, residual DCT/quant QF 50 -> int8 indices, reconstruction.
A GPU's share of the 240-frame clip at N GPUs is 240/N frames (GOPs are independent: sharding.frame_range), so one
GPU is timed on --frames frames (default 32 = its share at N = 8, rounded up to whole GOPs).  Prints one JSON object."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    H, W, T = 2160, 3840, a.frames
    clip_np = synth.clip(T, H, W, seed=21, margin=96)
    host = torch.from_numpy(clip_np).pin_memory()
    dev = host.cuda()
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=32, gop_len=4, qf=50.0,
                       metric=v.METRIC_WRAP8, static_thr=2000, coef_mode=v.COEF_I8_RINT)
    out = ce.alloc_device_outputs(T, want_coef=True, want_recon=True)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    for _ in range(2):
        ce.encode_device(dev, out, s)
    torch.cuda.synchronize()
    ce.ctx.enable_kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(a.iters):
        ce.encode_device(dev, out, s)
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    me, dct, calls = ce.ctx.kernel_times()
    ce.ctx.enable_kernel_timing(False)
    hout = ce.alloc_host_outputs(T, want_coef=True, want_recon=False, pinned=True)
    import time
    ce.encode_host(host, hout)
    t0 = time.perf_counter()
    for _ in range(a.iters):
        ce.encode_host(host, hout)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / a.iters * 1e3
    nP = ce.num_p_frames(T)
    print(json.dumps({"workload": "4K 2160x3840, bs 16, +/-32 full search (wrap8 + static test), DCT/quant QF50 int8, recon",
                      "frames": T, "p_frames": nP, "ms_per_clip_device": ms, "frames_per_s_device": T / ms * 1e3,
                      "me_ms": me / max(calls, 1), "dct_ms": dct / max(calls, 1),
                      "ms_per_clip_e2e": e2e_ms, "frames_per_s_e2e": T / e2e_ms * 1e3,
                      "same_results": bool(torch.equal(hout["mv"], out["mv"].cpu()) and torch.equal(hout["coef"], out["coef"].cpu())),
                      "projection": "GOPs are independent: N GPUs process N such shards concurrently (bench.py --gpus N measures the same weak scaling on config 2)"}))


if __name__ == "__main__":
    main()
