import os, sys, time
sys.path.insert(0, '/root/repo')
import torch, bench
import vcs_h264_b200 as v
clip = torch.from_numpy(bench.make_clip(1234)).pin_memory()
ce = v.ClipEncoder([bench.H, bench.W], block_size=bench.BS, search="full", search_range=bench.R, gop_len=bench.GOP,
                   qf=bench.QF, metric=0, static_thr=bench.STATIC_THR, coef_mode=v.COEF_I8_RINT)
hout = ce.alloc_host_outputs(bench.T, want_coef=True, want_recon=False, pinned=True)
for s in ("", "15", "1,1,2,3,3,2,2,1", "1"):
    if s: os.environ["VCS_PIPELINE_P"] = s
    else: os.environ.pop("VCS_PIPELINE_P", None)
    for _ in range(3): ce.encode_host(clip, hout)
    ce.ctx.enable_kernel_timing(True)
    n = 5
    t0 = time.perf_counter()
    for _ in range(n): ce.encode_host(clip, hout)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    me, dct, calls = ce.ctx.kernel_times()
    ce.ctx.enable_kernel_timing(False)
    print(f"sched {s or 'default':18s} wall {dt:7.3f} ms  sum ME {me/n:7.3f}  sum DCT {dct/n:6.3f}  launches/clip {calls/n:.0f}")
