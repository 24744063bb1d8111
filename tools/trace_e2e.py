"""Timeline of one vcs_encode_clip_host[_packed] call on the bench clip (VCS_TRACE=1: per segment upload done /
compute start / compute done / download done, ms from the start of the call).  python tools/trace_e2e.py [--dense]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import vcs_h264_b200 as v

dense = "--dense" in sys.argv
clip = torch.from_numpy(bench.make_clip(1234)).pin_memory()
ce = v.ClipEncoder([bench.H, bench.W], block_size=bench.BS, search="full", search_range=bench.R, gop_len=bench.GOP,
                   qf=bench.QF, metric=0, static_thr=bench.STATIC_THR, coef_mode=v.COEF_I8_RINT)
hout = ce.alloc_host_outputs(bench.T, want_coef=True, want_recon=False, pinned=True) if dense else ce.alloc_host_packed(bench.T, pinned=True)
if dense:
    del hout["cost"]
run = (lambda: ce.encode_host(clip, hout)) if dense else (lambda: ce.encode_host_packed(clip, hout))
for _ in range(3):
    run()
torch.cuda.synchronize()
os.environ["VCS_TRACE"] = "1"
t0 = time.perf_counter()
run()
torch.cuda.synchronize()
print(f"traced call: {(time.perf_counter() - t0) * 1e3:.3f} ms wall", flush=True)
