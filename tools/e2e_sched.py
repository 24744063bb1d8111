"""Times vcs_encode_clip_host[_packed] (workload C2) under different P-frame segment schedules (VCS_PIPELINE_P).
    python tools/e2e_sched.py [--dense] [schedule ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import vcs_h264_b200 as v

dense = "--dense" in sys.argv
scheds = [a for a in sys.argv[1:] if a != "--dense"] or ["", "3", "6", "1,2,3,4,6,8,8,8,4,1", "1,2,4,6,8,8,9,4,2,1", "1,1,2,3,4,6,8,8,8,3,1", "1,2,4,4,8,8,8,4,4,2", "1,3,4,8,8,8,8,4,1", ""]
clip = torch.from_numpy(bench.make_clip(1234)).pin_memory()
ce = v.ClipEncoder([bench.H, bench.W], block_size=bench.BS, search="full", search_range=bench.R, gop_len=bench.GOP,
                   qf=bench.QF, metric=0, static_thr=bench.STATIC_THR, coef_mode=v.COEF_I8_RINT)
hout = ce.alloc_host_outputs(bench.T, want_coef=True, want_recon=False, pinned=True) if dense else ce.alloc_host_packed(bench.T, pinned=True)
run = (lambda: ce.encode_host(clip, hout)) if dense else (lambda: ce.encode_host_packed(clip, hout))
ref = None
for s in scheds:
    if s:
        os.environ["VCS_PIPELINE_P"] = s
    else:
        os.environ.pop("VCS_PIPELINE_P", None)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    sig = (int(hout["mv"].to(torch.int64).sum()), int(hout["coef"].to(torch.int64).abs().sum()) if dense else (hout["lengths"], int(hout["nibbles"][:hout["lengths"][0]].to(torch.int64).sum())))
    ref = ref or sig
    print(f"sched {s or 'default':24s} {ms:7.3f} ms  {bench.T / ms * 1e3:7.1f} fps  same={sig == ref}", flush=True)
