"""Run the tiled ME kernel alone (SAD and wrap8) on a 1080p clip (profiling helper)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcs_h264_b200 as v
from vcs_h264_b200 import synth
T, H, W, R = int(os.environ.get("T", 24)), 1080, 1920, int(os.environ.get("R", 16))
clip = torch.from_numpy(synth.clip(T, H, W, seed=1)).cuda()
for metric in (v.METRIC_SAD, v.METRIC_WRAP8):
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=R, gop_len=4, metric=metric)
    out = ce.alloc_device_outputs(T, want_coef=False, want_recon=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0.record(); ce.me_device(clip, out); e1.record(); torch.cuda.synchronize()
    nP = ce.num_p_frames(T)
    print("R", R, "metric", metric, "ME ms for", nP, "P-frames:", e0.elapsed_time(e1), "us/frame", 1e3 * e0.elapsed_time(e1) / nP)
