"""Run one tiled-ME configuration vs the oracle (debug helper, one case per process)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcs_h264_b200 as v
from oracle import oracle as orc
from vcs_h264_b200 import synth

def run(bs, lo, hi, slack, H, W, metric, thr, kernel=2):
    c = v._capi
    clip = synth.clip(2, H, W, seed=bs + hi, margin=48)
    cur, ref = clip[1].copy(), clip[0]
    cur[:bs, :2 * bs] = ref[:bs, :2 * bs]
    p = c.me_reference_params(H, W, bs)
    p.lo, p.hi, p.step, p.slack, p.metric, p.static_thr, p.kernel = lo, hi, 1, slack, metric, thr, kernel
    N = c.num_blocks(H, W, bs)
    mv = np.empty((N, 2), np.int16); cost = np.empty(N, np.uint32); fl = np.empty(N, np.uint8)
    ctx = v.runtime.get_context()
    ctx.call("vcs_me_search_host", p, cur.ctypes.data, ref.ctypes.data, mv.ctypes.data, cost.ctypes.data, fl.ctypes.data)
    omv, ocost, ofl = orc.me(cur, ref, bs, lo, hi, 1, slack, metric=metric, static_thr=thr)
    ok = np.array_equal(mv.astype(np.int32), omv) and np.array_equal(cost, ocost) and np.array_equal(fl, ofl)
    bad = int((mv.astype(np.int32) != omv).any(1).sum())
    return ok, bad, N

if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]]
    try:
        print("CASE", args, run(*args))
    except Exception as e:
        print("CASE", args, "ERROR", str(e)[-200:])
