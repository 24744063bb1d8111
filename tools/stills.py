#!/usr/bin/env python
"""BASELINE.json config 4 (DCT part): 8x8 DCT / quantise / reconstruct of synthetic 4K stills
(images/bigImg.png is missing from the reference repo) at quality 10 / 50 / 99, the
arithmetic of DCTCompression/dct.py:169-208.  Prints one JSON object."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import _capi, synth
    H, W, B = 2160, 3840, 16                       # a batch of 16 stills per launch
    stills = torch.from_numpy(np.stack([synth.still(H, W, seed=100 + k) for k in range(2)])).cuda()
    stills = stills.repeat(B // 2, 1, 1, 1).contiguous()
    ctx = v.runtime.get_context(0)
    s = torch.cuda.current_stream()
    ctx.set_stream(s.cuda_stream)
    rows = []
    idx = torch.empty((B, 3, H, W), dtype=torch.int16, device="cuda")
    rec = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
    for qf in (10.0, 50.0, 99.0):
        ctx.set_q(_capi.q_tables(qf))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = {}
        for name, fn in (("compress", lambda k: ctx.call("vcs_compress_dev", H, W, _capi.ptr(stills[k]), 2, _capi.ptr(idx[k]))),
                         ("decompress", lambda k: ctx.call("vcs_decompress_dev", H, W, 2, _capi.ptr(idx[k]), None, _capi.ptr(rec[k])))):
            for k in range(B):
                fn(k)
            torch.cuda.synchronize()
            e0.record()
            for k in range(B):
                fn(k)
            e1.record()
            torch.cuda.synchronize()
            times[name] = e0.elapsed_time(e1) / B
        sp = v.sparsity_device(ctx, idx, 2)
        err = (rec[0].float() - stills[0].float())
        psnr = float(10 * torch.log10(255.0 ** 2 / (err ** 2).mean()))
        px = H * W
        rows.append({"qf": qf, "compress_ms_per_still": times["compress"], "decompress_ms_per_still": times["decompress"],
                     "compress_GBps": px * (3 + 6) / (times["compress"] * 1e-3) / 1e9,
                     "decompress_GBps": px * (6 + 3) / (times["decompress"] * 1e-3) / 1e9,
                     "sparsity": sp, "psnr_db": psnr})
    # intra mode decision on the same 4K still (IntraframeCompression/intraframe.py:24-317)
    # any uint8 planes will do for timing: the still's own B, G, R channels stand in for Y, Cr, Cb
    Y, Cr, Cb = (stills[0][..., k].contiguous() for k in range(3))
    o = [torch.empty((H, W), dtype=torch.int32, device="cuda") for _ in range(4)]
    m4 = torch.empty((H // 4, W // 4), dtype=torch.uint8, device="cuda")
    intra = {}
    for name, fn in (("luma4x4", lambda: ctx.call("vcs_intra_luma4x4_dev", H, W, _capi.ptr(Y), _capi.ptr(o[0]), _capi.ptr(o[1]), _capi.ptr(m4))),
                     ("luma16x16", lambda: ctx.call("vcs_intra_luma16x16_dev", H, W, _capi.ptr(Y), _capi.ptr(o[0]), _capi.ptr(o[1]), _capi.ptr(m4))),
                     ("chroma8x8", lambda: ctx.call("vcs_intra_chroma8x8_dev", H, W, _capi.ptr(Cr), _capi.ptr(Cb), _capi.ptr(o[0]), _capi.ptr(o[1]), _capi.ptr(o[2]), _capi.ptr(o[3]), _capi.ptr(m4)))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        intra[name + "_ms"] = e0.elapsed_time(e1) / 10
    # 4:2:0 chroma subsampling of the same still (ChromaSubsampling/chroma.py:9-41), device resident
    h2, w2 = (H + 1) // 2, (W + 1) // 2
    Yp = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    crs, cbs = (torch.empty((h2, w2), dtype=torch.uint8, device="cuda") for _ in range(2))
    back = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    chroma = {}
    for name, fn, nbytes in (("subsample", lambda: ctx.call("vcs_chroma420_dev", H, W, _capi.ptr(stills[0]), _capi.ptr(Yp), _capi.ptr(crs), _capi.ptr(cbs)), H * W * 4 + 2 * h2 * w2),
                             ("to_bgr", lambda: ctx.call("vcs_chroma420_to_bgr_dev", H, W, _capi.ptr(Yp), _capi.ptr(crs), _capi.ptr(cbs), _capi.ptr(back)), H * W * 4 + 2 * h2 * w2)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        chroma[name + "_ms"] = ms
        chroma[name + "_GBps"] = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": "synthetic 2160x3840 stills, 8x8 DCT f64, rint quantiser -> int16", "rows": rows,
                      "intra_4k": intra, "chroma420_4k": chroma}))


if __name__ == "__main__":
    main()
