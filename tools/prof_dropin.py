import cProfile, pstats, sys, os, io, contextlib
sys.path.insert(0, '/root/repo')
import torch
from vcs_h264_b200 import main as drv, synth
frames = list(synth.clip(40, 360, 640, seed=5))
with contextlib.redirect_stdout(io.StringIO()):
    drv.run(frames, block_size=8, mode="frame")
pr = cProfile.Profile()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    drv.run(frames, block_size=8, mode="frame")
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:4500])
