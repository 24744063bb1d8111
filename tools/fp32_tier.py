#!/usr/bin/env python
"""The fp32 tier of the DCT stage against the exact float64 tier on the bench clip (workload C2, int8 indices +
reconstruction): flip counters, PSNR and the DCT-stage time of both.  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import vcs_h264_b200 as v
from vcs_h264_b200 import _capi

clip = torch.from_numpy(bench.make_clip(1234)).cuda()
res, outs = {}, {}
for bits in (64, 32):
    ce = v.ClipEncoder([bench.H, bench.W], block_size=bench.BS, search="full", search_range=bench.R, gop_len=bench.GOP,
                       qf=bench.QF, metric=0, static_thr=bench.STATIC_THR, coef_mode=v.COEF_I8_RINT, dct_precision=bits)
    out = ce.alloc_device_outputs(bench.T)
    ce.encode_device(clip, out)
    torch.cuda.synchronize()
    s = torch.cuda.current_stream()
    ce.ctx.set_stream(s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(5):
        e0.record()
        ce.ctx.call("vcs_residual_dct_clip_dev", bench.H, bench.W, bench.BS, _capi.ptr(clip), bench.T, bench.GOP,
                    _capi.ptr(out["mv"]), v.COEF_I8_RINT, _capi.ptr(out["coef"]), _capi.ptr(out["recon"]))
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ce.ctx.use_own_stream()
    res[f"dct_stage_ms_fp{bits}"] = min(ms)
    outs[bits] = (ce, out)
(ce64, a), (_, b) = outs[64], outs[32]
fc = v.flip_counters(ce64.ctx, v.COEF_I8_RINT, a["coef"], b["coef"], a["recon"], b["recon"])
d = (a["recon"].to(torch.int16) - b["recon"].to(torch.int16)).abs()
res.update(fc)
res["index_flip_fraction"] = fc["index_flips"] / fc["indices"]
res["pixel_flip_fraction"] = fc["pixel_flips"] / fc["pixels"]
res["max_abs_pixel_difference"] = int(d.max())
res["motion_vectors_equal"] = bool(torch.equal(a["mv"], b["mv"]))
res["workload"] = "C2 bench clip, 45 P-frames 1080p, int8 indices QF 50 + reconstruction; fp32 tier against the exact tier"
print(json.dumps(res))
