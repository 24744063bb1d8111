"""Run the residual/DCT stage alone on a 1080p clip with the bench's settings (int8 indices + reconstruction).
DCT_BITS=32 times the fp32 tier instead of the exact float64 one."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcs_h264_b200 as v
from vcs_h264_b200 import synth, _capi
T, H, W = 60, 1080, 1920
clip = torch.from_numpy(synth.clip(T, H, W, seed=1)).cuda()
ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=16, gop_len=4, coef_mode=v.COEF_I8_RINT,
                   dct_precision=int(os.environ.get("DCT_BITS", "64")))
out = ce.alloc_device_outputs(T)
ce.encode_device(clip, out)
torch.cuda.synchronize()
ctx = ce.ctx
s = torch.cuda.current_stream()
ctx.set_stream(s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):
    e0.record()
    ctx.call("vcs_residual_dct_clip_dev", H, W, 16, _capi.ptr(clip), T, 4, _capi.ptr(out["mv"]), v.COEF_I8_RINT,
             _capi.ptr(out["coef"]), _capi.ptr(out["recon"]))
    e1.record(); torch.cuda.synchronize()
    print("dct stage ms for", ce.num_p_frames(T), "P-frames:", e0.elapsed_time(e1))
