"""Whole-frame parity on the configurations BASELINE.json names (no strips, no sampling).

  C2  the bench clip itself -- synth.clip(60, 1080, 1920, 1234), +/-16, bs 16, reference cost + static test, QF 50 --
      through ClipEncoder exactly as bench.py builds it: ALL 45 P-frames, ALL 8040 macroblocks, every int8 index and
      every reconstructed pixel against the CPU oracle;
  C3  one whole 2160x3840 +/-32 P-frame (32 400 macroblocks x 4225 candidates) against the oracle;
  C1  the reference's own default input, videos/traffic_cut.mp4 (main.py:13-16), all 152 frames through the drop-in
      Encoder on the CUDA path, against sha256 pins taken from the UNMODIFIED reference
      (tests/golden/make_golden_traffic.py): the 114 x 3600 motion vectors, the static-block counts, and the float64
      coefficient planes and decoded frames of P-frames 1 and 35.
"""
import contextlib
import hashlib
import io
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vcs():
    import vcs_h264_b200 as v
    v.runtime.get_context()          # fails loudly without a GPU / the built extension
    return v


def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_c2_bench_clip_every_macroblock(vcs, orc):
    import torch
    from vcs_h264_b200 import synth
    T, H, W, bs, R, gop = 60, 1080, 1920, 16, 16, 4
    clip = synth.clip(T, H, W, seed=1234)
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=gop, qf=50.0,
                         metric=vcs.METRIC_WRAP8, static_thr=2000, coef_mode=vcs.COEF_I8_RINT, device=0)
    dev_in = torch.from_numpy(clip).cuda(0)
    dout = ce.alloc_device_outputs(T, want_coef=True, want_recon=True)
    ce.encode_device(dev_in, dout)
    torch.cuda.synchronize()
    hout = ce.encode_host(clip, want_coef=True, want_recon=True)       # the pipelined host path (p_off segments)
    d = {k: v.cpu().numpy() for k, v in dout.items()}
    prm = orc.symmetric_search_params(R)
    Q = orc.qtables(50.0)
    nthreads = _cores()
    bad = []
    for p, t in enumerate(ce.p_frame_indices(T)):
        o = orc.encode_p(clip[t], clip[(t // gop) * gop], bs, metric=orc.METRIC_WRAP8, static_thr=2000, Q=Q,
                         round_mode=1, simd=True, nthreads=nthreads, **prm)
        for name, res in (("device", d), ("host", {k: np.asarray(v) for k, v in hout.items()})):
            m = dict(mv=int((res["mv"][p].astype(np.int32) != o["mv"]).any(1).sum()),
                     cost=int((res["cost"][p].view(np.uint32) != o["cost"]).sum()),
                     flags=int((res["flags"][p] != o["flags"]).sum()),
                     index=int((res["coef"][p].astype(np.float64) != o["planes"]).sum()),
                     pixel=int((res["recon"][p] != o["recon"]).sum()))
            if any(m.values()):
                bad.append((p, name, m))
    assert not bad, bad[:5]


def test_c3_whole_4k_frame(vcs, orc):
    from vcs_h264_b200 import synth
    H, W, bs, R = 2160, 3840, 16, 32
    clip = synth.clip(2, H, W, seed=77)
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=2, qf=50.0,
                         metric=vcs.METRIC_WRAP8, static_thr=2000, coef_mode=vcs.COEF_I8_RINT, device=0)
    out = ce.encode_host(clip, want_coef=True, want_recon=True)
    o = orc.encode_p(clip[1], clip[0], bs, metric=orc.METRIC_WRAP8, static_thr=2000, Q=orc.qtables(50.0),
                     round_mode=1, simd=True, nthreads=_cores(), **orc.symmetric_search_params(R))
    assert np.array_equal(np.asarray(out["mv"][0]).astype(np.int32), o["mv"])
    assert np.array_equal(np.asarray(out["cost"][0]).view(np.uint32), o["cost"])
    assert np.array_equal(np.asarray(out["flags"][0]), o["flags"])
    assert np.array_equal(np.asarray(out["coef"][0]).astype(np.float64), o["planes"])
    assert np.array_equal(np.asarray(out["recon"][0]), o["recon"])


def test_c1_traffic_cut_all_frames(vcs):
    cv2 = pytest.importorskip("cv2")
    with open(os.path.join(ROOT, "tests", "golden", "golden_traffic_meta.json")) as f:
        meta = json.load(f)
    path = os.path.join(ROOT, "tests", "golden", meta["file"])
    assert hashlib.sha256(open(path, "rb").read()).hexdigest() == meta["file_sha256"]
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    cap.release()
    assert len(frames) == meta["frames"] and list(frames[0].shape[:2]) == meta["shape"]
    if _sha(np.stack(frames)) != meta["frames_sha16"]:
        pytest.skip("this host's video decoder yields different pixels than the one the pins were taken with")
    H, W = meta["shape"]
    enc = vcs.Encoder(["I", "P", "P", "P"], [H, W], 8, True)           # main.py:13-16,29-30
    with contextlib.redirect_stdout(io.StringIO()):
        for n, f in enumerate(frames):                                 # main.py:34-41
            enc.encode_frame(f, n)
    P = [f for f in enc.encoded_frames if f.t == "P"]
    assert len(P) == meta["n_p"] == 114
    mv = np.asarray([f.mv for f in P], np.int32)
    assert mv.shape == (114, 3600, 2)
    assert _sha(mv) == meta["mv_sha16"] == "ddae5b429d9f442e"
    statics = [int(((m[:, 0] == 0) & (m[:, 1] == 0)).sum()) for m in mv]
    lo, hi, mean = meta["static_minmaxmean"]
    assert (min(statics), max(statics)) == (lo, hi) == (3328, 3597) and abs(np.mean(statics) - mean) < 1e-9
    assert _sha(np.asarray(P[0].c, np.int32)) == meta["coords_sha16"]
    dec = vcs.Decoder(enc.encoded_frames, 25, [H, W], enc.ref_frames, 8, True)
    for n in (1, 35):
        fr = enc.encoded_frames[n]
        assert fr.t == "P" and fr.i == n
        assert _sha(np.stack(fr.r)) == meta[f"frame{n}_planes_sha16"], f"float64 planes of frame {n}"
        with contextlib.redirect_stdout(io.StringIO()):
            final = dec._reconstruct_P_frame(fr, True)
        assert _sha(final) == meta[f"frame{n}_final_sha16"], f"decoded frame {n}"
