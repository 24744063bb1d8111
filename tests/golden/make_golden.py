#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference is imported as-is (matplotlib is absent, so a MagicMock stands in for it; the hot
path never calls it).  Generalised settings use only attributes the reference exposes:
`mp.search_window_size = R` (motion.py:18) and the module-global name `round` that
motion.py:132 resolves at call time (`motion.round = lambda x: 1` -> step 1).

Nothing here is product code; the outputs pin oracle/vcs_oracle.c (tests/test_oracle_golden.py)
and, through it, the CUDA path (tests/test_gpu_*.py).
"""
import contextlib
import hashlib
import io
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

sys.modules.setdefault("matplotlib", MagicMock())
sys.modules.setdefault("matplotlib.pyplot", MagicMock())
sys.path.insert(0, os.path.join(REF, "InterframeCompression"))
import cv2  # noqa: E402
import motion as ref_motion  # noqa: E402
from motion import MotionProcessor  # noqa: E402
from DCTcompressor import DCTCompressor, QY, QC  # noqa: E402
from encoder import Encoder  # noqa: E402
from decoder import Decoder  # noqa: E402


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def read_video(path, limit=None):
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok or (limit and len(frames) >= limit):
            break
        frames.append(f)
    cap.release()
    return frames


def ref_me(cur, ref, bs, R=None, step1=False):
    """process_motion_prediction on the unmodified class; returns mv, coords, cost, static."""
    H, W = cur.shape[:2]
    mp = MotionProcessor(bs, [H, W])
    if R is not None:
        mp.search_window_size = R
    if step1:
        ref_motion.round = lambda x: 1
    try:
        with quiet():
            mvs, coords = mp.process_motion_prediction(cur, ref)
            # per-MB winning cost: recompute from _find_match's second return value
            blocks, _ = mp._split_frame_into_mblocks(cur)
            cost = np.zeros(len(blocks), np.int64)
            static = np.zeros(len(blocks), np.uint8)
            for k, (blk, c) in enumerate(zip(blocks, coords)):
                at = ref[c[1]:c[1] + bs, c[0]:c[0] + bs]
                S = int(np.sum(np.abs(cv2.subtract(at, blk))))
                if S <= ref_motion.SIMILARITY_THRESHOLD:
                    static[k] = 1
                    cost[k] = S
                else:
                    best_coord, best_block = mp._find_match(ref, blk, c)
                    if best_block.shape == blk.shape:
                        cost[k] = int(np.sum(np.abs(best_block - blk)))
                    else:
                        cost[k] = -1   # no candidate
    finally:
        if step1:
            del ref_motion.round
    return (np.array(mvs, np.int32).reshape(-1, 2), np.array(coords, np.int32).reshape(-1, 2),
            cost, static)


def ref_p_frame(cur, ref, bs_me, Q=None):
    """Encoder._process_P_frame + Decoder._reconstruct_P_frame with ME bs decoupled from the
    8x8 DCT (the reference's Encoder passes one block size to both)."""
    H, W = cur.shape[:2]
    mp = MotionProcessor(bs_me, [H, W])
    dc = DCTCompressor(8)
    if Q is not None:
        dc.Q = Q
    with quiet():
        mvs, coords = mp.process_motion_prediction(cur, ref)
        pred = mp.reconstruct_from_motion_vectors(mvs, ref, coords)
        resid = mp.get_residuals(input_frame=cur, reconstructed=pred)
        planes = dc.compress(resid)
        dec = dc.decompress(compressed=planes, imshape=pred.shape)
        final = pred + dec
        planes_r = [np.round(p) for p in planes]
        dec_r = dc.decompress(compressed=planes_r, imshape=pred.shape)
        final_r = pred + dec_r
    return dict(mv=np.array(mvs, np.int32).reshape(-1, 2), pred=pred, resid=resid,
                planes=np.stack(planes), dec=dec, final=final,
                planes_r=np.stack(planes_r), dec_r=dec_r, final_r=final_r)


def q_for(qf):
    """The module-level Q expression of DCTcompressor.py:29-38 / dct.py:157-166 for another QF."""
    if qf < 50 and qf > 1:
        scale = 50 / qf
    elif qf < 100:
        scale = (100 - qf) / 50
    else:
        raise ValueError
    return [np.clip(np.round(QY * scale), 1, 255), np.clip(np.round(QC * scale), 1, 255),
            np.clip(np.round(QC * scale), 1, 255)]


def synth_pair(H, W, seed, shift=(3, -2), noise=2):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H + 64, W + 64, 3), dtype=np.uint8)
    base = cv2.GaussianBlur(base, (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    ref = base[32:32 + H, 32:32 + W].copy()
    cur = base[32 + shift[1]:32 + shift[1] + H, 32 + shift[0]:32 + shift[0] + W].astype(np.int16)
    cur = np.clip(cur + rng.integers(-noise, noise + 1, cur.shape), 0, 255).astype(np.uint8)
    return cur, ref


def main():
    out = {}
    meta = {"numpy": np.__version__, "cv2": cv2.__version__}

    # ---- 1. DCT primitives -----------------------------------------------------------------
    dc = DCTCompressor(8)
    rng = np.random.default_rng(7)
    out["dctmat"] = dc._dctMatrix()
    blocks = rng.integers(-128, 128, (64, 8, 8)).astype(np.int16)
    blocks[0] = 1 - 128 + 127  # flat block of 0
    blocks[1] = 1              # all ones: DC/16 lands next to .5 (SURVEY 7 "DCT parity")
    blocks[2] = 3
    blocks[3] = -128
    blocks[4] = 127
    out["dct_in"] = blocks
    out["dct_out"] = np.stack([dc._dct2(b) for b in blocks])
    coefs = rng.normal(0, 200, (64, 8, 8))
    coefs[0] = 0
    out["idct_in"] = coefs
    out["idct_out"] = np.stack([dc._idct2(c) for c in coefs])
    for qf in (1.0, 10.0, 49.0, 50.0, 75.0, 99.0):
        out[f"Q_{int(qf)}"] = np.stack(q_for(qf))
    from DCTcompressor import Q as Qmod
    assert all(np.array_equal(a, b) for a, b in zip(Qmod, q_for(50.0)))
    out["Q_module"] = np.stack(Qmod)

    # float64 -> uint8 store semantics (DCTcompressor.py:81,88) and astype (dct.py:204)
    castv = np.array([-300.7, -256.0, -255.9, -129.2, -128.0, -1.5, -1.0, -0.9999999, -0.5, -1e-13,
                      0.0, 1e-13, 0.5, 0.9999999999, 1.0, 127.99, 128.0, 255.0, 255.9, 256.0,
                      300.7, 511.5, 1000.25, -1000.25], np.float64)
    tgt = np.zeros(castv.shape, "uint8")
    tgt[:] = castv
    out["cast_in"] = castv
    out["cast_setitem"] = tgt
    with np.errstate(all="ignore"):
        out["cast_astype"] = castv.astype(np.uint8)

    # ---- 2. colour conversion ----------------------------------------------------------------
    cols = rng.integers(0, 256, (4096, 1, 3), dtype=np.uint8)
    out["bgr_in"] = cols
    out["ycrcb_out"] = cv2.cvtColor(cols, cv2.COLOR_BGR2YCR_CB)
    out["ycrcb_in"] = cols
    out["bgr_out"] = cv2.cvtColor(cols, cv2.COLOR_YCR_CB2BGR)
    # exhaustive check of the restated formulas against cv2 (result recorded in meta)
    from oracle import oracle as orc
    allc = np.stack(np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij"),
                    -1).astype(np.uint8).reshape(-1, 1, 3)
    bad_f = bad_i = 0
    for k in range(0, allc.shape[0], 1 << 20):
        chunk = np.ascontiguousarray(allc[k:k + (1 << 20)])
        bad_f += int((cv2.cvtColor(chunk, cv2.COLOR_BGR2YCR_CB) != orc.bgr2ycrcb(chunk)).sum())
        bad_i += int((cv2.cvtColor(chunk, cv2.COLOR_YCR_CB2BGR) != orc.ycrcb2bgr(chunk)).sum())
    meta["colour_exhaustive_mismatches"] = {"bgr2ycrcb": bad_f, "ycrcb2bgr": bad_i}
    print("colour exhaustive mismatches", bad_f, bad_i)

    # ---- 3. ME on synthetic / adversarial frames ----------------------------------------------
    me_cases = []

    def add_me(name, cur, ref, bs, R=None, step1=False):
        mv, coords, cost, static = ref_me(cur, ref, bs, R, step1)
        out[f"me_{name}_cur"] = cur
        out[f"me_{name}_ref"] = ref
        out[f"me_{name}_mv"] = mv
        out[f"me_{name}_coords"] = coords
        out[f"me_{name}_cost"] = cost
        out[f"me_{name}_static"] = static
        me_cases.append(dict(name=name, bs=bs, R=(2 * bs if R is None else R), step1=step1))

    for bs in (4, 8, 16):
        cur, ref = synth_pair(72, 104, 100 + bs, shift=(3, -2))
        add_me(f"synth_bs{bs}", cur, ref, bs)
    cur, ref = synth_pair(53, 77, 5, shift=(-4, 5))      # non-multiple sizes, partial MBs dropped
    add_me("ragged_bs8", cur, ref, 8)
    cur, ref = synth_pair(48, 64, 6, shift=(2, 1))
    add_me("step1_bs8_R16", cur, ref, 8, R=16, step1=True)  # literal step-1 oracle
    cur, ref = synth_pair(64, 96, 8, shift=(-5, 3))
    add_me("step1_bs16_R32", cur, ref, 16, R=32, step1=True)
    cur, ref = synth_pair(40, 56, 9, shift=(1, 1))
    add_me("R8_bs8", cur, ref, 8, R=8)                   # window smaller than default
    # ties: flat frames -> every candidate costs the same; first in scan order must win
    flat_ref = np.full((40, 48, 3), 50, np.uint8)
    flat_cur = np.full((40, 48, 3), 90, np.uint8)       # ref-cur wraps: cost 216/byte, not static?
    add_me("ties_flat", flat_cur, flat_ref, 8)
    flat_cur2 = np.full((40, 48, 3), 10, np.uint8)      # ref>cur: one-sided sum 40*192 > 2000
    add_me("ties_flat2", flat_cur2, flat_ref, 8)
    # periodic texture -> many exact ties at different offsets
    yy, xx = np.mgrid[0:48, 0:64]
    per = (((xx % 6) * 40 + (yy % 3) * 13) % 256).astype(np.uint8)
    per_ref = np.stack([per, per // 2, 255 - per], -1)
    per_cur = np.roll(per_ref, (3, 6), (0, 1))
    per_cur[::5, ::7] ^= 0x55
    add_me("ties_periodic", per_cur, per_ref, 8)
    # all static (identical frames) and nearly static
    cur, ref = synth_pair(32, 48, 11, shift=(0, 0), noise=0)
    add_me("all_static", cur, ref, 8)
    cur, ref = synth_pair(32, 48, 12, shift=(0, 0), noise=3)
    add_me("noise_static", cur, ref, 8)
    # tiny frames: zero candidates (H == bs), one MB only
    cur, ref = synth_pair(8, 8, 13, shift=(1, 1))
    add_me("tiny_8x8", cur, ref, 8)
    cur, ref = synth_pair(16, 40, 14, shift=(2, 0))
    add_me("tiny_16x40_bs16", cur, ref, 16)
    cur, ref = synth_pair(9, 30, 15, shift=(2, 0))
    add_me("thin_9x30_bs8", cur, ref, 8)
    # extremes of the wrap metric
    rng2 = np.random.default_rng(21)
    cur = rng2.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    ref = rng2.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    add_me("random_bs8", cur, ref, 8)
    add_me("random_bs16", cur, ref, 16)
    add_me("random_bs4", cur[:24, :32].copy(), ref[:24, :32].copy(), 4)
    meta["me_cases"] = me_cases

    # ---- 4. real clip: crops of traffic_cut + full-size summaries -----------------------------
    frames = read_video(os.path.join(REF, "videos", "traffic_cut.mp4"))
    meta["traffic_cut_frames"] = len(frames)
    meta["traffic_cut_shape"] = list(frames[0].shape)
    # moving region crop (cars) so that non-static blocks exist; multiples of 16
    f0, f1, f3 = frames[0], frames[1], frames[3]
    y0, x0, ch, cw = 168, 256, 96, 160
    crop = lambda f: np.ascontiguousarray(f[y0:y0 + ch, x0:x0 + cw])
    add_me("traffic_crop_bs8", crop(f3), crop(f0), 8)
    add_me("traffic_crop_bs16", crop(f3), crop(f0), 16)

    pf_cases = []
    for name, cur, ref, bs in (("traffic_crop_bs8", crop(f3), crop(f0), 8),
                               ("traffic_crop_bs16", crop(f1), crop(f0), 16),
                               ("synth_bs16", *synth_pair(64, 96, 31, shift=(4, -3)), 16)):
        r = ref_p_frame(cur, ref, bs)
        out[f"pf_{name}_cur"] = cur
        out[f"pf_{name}_ref"] = ref
        for k, v in r.items():
            out[f"pf_{name}_{k}"] = v
        pf_cases.append(dict(name=name, bs=bs, qf=50))
    # stills path (dct.py): compress/decompress of an image at QF 10/50/99 with rounding
    still = crop(frames[10])
    out["still_img"] = still
    for qf in (10.0, 50.0, 99.0):
        d = DCTCompressor(8)
        d.Q = q_for(qf)
        with quiet():
            pl = d.compress(still)
            plr = [np.round(p) for p in pl]
            dec = d.decompress(plr, still.shape)
        out[f"still_q{int(qf)}_planes"] = np.stack(pl)
        out[f"still_q{int(qf)}_dec"] = dec
        nz = sum(np.count_nonzero(p) for p in plr)
        meta[f"still_q{int(qf)}_sparsity"] = 1.0 - nz / float(sum(p.size for p in plr))
    meta["pf_cases"] = pf_cases

    # full-size: every P-frame of traffic_cut through the literal reference (bs 8, I-P-P-P)
    H, W = frames[0].shape[:2]
    all_mv = []
    statics = []
    for n, f in enumerate(frames):
        if n % 4 == 0:
            continue
        mv, coords, cost, static = ref_me(f, frames[(n // 4) * 4], 8)
        all_mv.append(mv)
        statics.append(int(static.sum()))
    all_mv = np.stack(all_mv).astype(np.int32)
    meta["traffic_full_mv_sha16"] = hashlib.sha256(all_mv.tobytes()).hexdigest()[:16]
    meta["traffic_full_static_minmaxmean"] = [min(statics), max(statics), float(np.mean(statics))]
    # the oracle on the same frames (recorded so the CPU test can re-check without the video)
    all_o = []
    p = orc.reference_search_params(8)
    for n, f in enumerate(frames):
        if n % 4 == 0:
            continue
        mv, _, _ = orc.me(f, frames[(n // 4) * 4], 8, **p)
        all_o.append(mv)
    meta["traffic_full_oracle_mv_sha16"] = hashlib.sha256(np.stack(all_o).tobytes()).hexdigest()[:16]
    print("traffic full sha ref/oracle", meta["traffic_full_mv_sha16"],
          meta["traffic_full_oracle_mv_sha16"])

    # full P-frame (frame 1 and 35) through reference vs oracle: record mismatch counts
    for n in (1, 35):
        r = ref_p_frame(frames[n], frames[(n // 4) * 4], 8)
        o = orc.encode_p(frames[n], frames[(n // 4) * 4], 8, **p)
        o_r = orc.encode_p(frames[n], frames[(n // 4) * 4], 8, round_mode=1, **p)
        meta[f"traffic_frame{n}_mismatch"] = dict(
            mv=int((r["mv"] != o["mv"]).sum()), planes=int((r["planes"] != o["planes"]).sum()),
            final=int((r["final"] != o["recon"]).sum()),
            planes_r=int((r["planes_r"] != o_r["planes"]).sum()),
            final_r=int((r["final_r"] != o_r["recon"]).sum()))
        print("frame", n, meta[f"traffic_frame{n}_mismatch"])

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.npz"),
          os.path.getsize(os.path.join(HERE, "golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
