"""Pins oracle/vcs_oracle.c against vectors produced by the unmodified reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest


def test_simd_costs_match_scalar(orc):
    assert orc.selfcheck_simd() == 0


def test_dct_matrix_bit_exact(orc, golden):
    # DCTcompressor.py:124-133
    assert np.array_equal(orc.dct_matrix(), golden["dctmat"])


@pytest.mark.parametrize("qf", [1, 10, 49, 50, 75, 99])
def test_qtables(orc, golden, qf):
    # DCTcompressor.py:29-38
    assert np.array_equal(orc.qtables(float(qf)), golden[f"Q_{qf}"])


def test_qtables_module_default_and_invalid(orc, golden):
    assert np.array_equal(orc.qtables(50.0), golden["Q_module"])
    with pytest.raises(ValueError):
        orc.qtables(100.0)


def test_dct2_idct2_bit_exact(orc, golden):
    # DCTcompressor.py:111-121: np.matmul chains == sequential-k fma
    for x, d in zip(golden["dct_in"], golden["dct_out"]):
        assert np.array_equal(orc.dct2(x), d)
    for x, d in zip(golden["idct_in"], golden["idct_out"]):
        assert np.array_equal(orc.idct2(x), d)


def test_colour_conversion(orc, golden):
    assert np.array_equal(orc.bgr2ycrcb(golden["bgr_in"]), golden["ycrcb_out"])
    assert np.array_equal(orc.ycrcb2bgr(golden["ycrcb_in"]), golden["bgr_out"])


def test_colour_exhaustive_was_clean(golden_meta):
    assert golden_meta["colour_exhaustive_mismatches"] == {"bgr2ycrcb": 0, "ycrcb2bgr": 0}


def test_uint8_cast_semantics(golden):
    # DCTcompressor.py:81,88 -- (uint8)(int64)trunc(x) for every value the IDCT can produce
    v = golden["cast_in"]
    want = golden["cast_setitem"]
    got = (np.trunc(v).astype(np.int64) & 0xFF).astype(np.uint8)
    inr = np.abs(v) < 2 ** 31
    assert np.array_equal(got[inr], want[inr])


def _me_params(orc, case):
    if case["step1"]:
        return orc.reference_search_params(case["bs"], R=case["R"], step=1)
    return orc.reference_search_params(case["bs"], R=case["R"])


def test_me_cases(orc, golden, golden_meta):
    """MotionProcessor.process_motion_prediction (motion.py:20-36) on every golden case:
    MVs, coords, static flags and winning cost."""
    assert len(golden_meta["me_cases"]) >= 20
    for case in golden_meta["me_cases"]:
        n = case["name"]
        cur, ref = golden[f"me_{n}_cur"], golden[f"me_{n}_ref"]
        for simd in (False, True):
            mv, cost, flags = orc.me(cur, ref, case["bs"], simd=simd, **_me_params(orc, case))
            assert np.array_equal(mv, golden[f"me_{n}_mv"]), n
            assert np.array_equal(flags & 1, golden[f"me_{n}_static"]), n
            gc = golden[f"me_{n}_cost"]
            ok = gc >= 0
            assert np.array_equal(cost[ok].astype(np.int64), gc[ok]), n
            assert np.all(flags[~ok] == 2), n
        H, W = cur.shape[:2]
        assert np.array_equal(orc.block_coords(H, W, case["bs"]), golden[f"me_{n}_coords"]), n


def test_p_frame_pipeline(orc, golden, golden_meta):
    """encoder.py:49-70 + decoder.py:52-69 stage by stage, un-rounded (inter path) and
    rounded (dct.py:179) variants, bit-exact including float64 planes."""
    for case in golden_meta["pf_cases"]:
        n, bs = case["name"], case["bs"]
        cur, ref = golden[f"pf_{n}_cur"], golden[f"pf_{n}_ref"]
        p = orc.reference_search_params(bs)
        mv, _, _ = orc.me(cur, ref, bs, **p)
        assert np.array_equal(mv, golden[f"pf_{n}_mv"])
        pred = orc.mc(ref, bs, mv)
        assert np.array_equal(pred, golden[f"pf_{n}_pred"])
        resid = orc.residual(cur, pred)
        assert np.array_equal(resid, golden[f"pf_{n}_resid"])
        planes = orc.compress(resid)
        assert np.array_equal(planes, golden[f"pf_{n}_planes"])
        dec = orc.decompress(planes)
        assert np.array_equal(dec, golden[f"pf_{n}_dec"])
        assert np.array_equal(orc.add_wrap(pred, dec), golden[f"pf_{n}_final"])
        planes_r = orc.compress(resid, round_mode=1)
        assert np.array_equal(planes_r, golden[f"pf_{n}_planes_r"])
        assert np.array_equal(orc.decompress(planes_r), golden[f"pf_{n}_dec_r"])
        fused = orc.encode_p(cur, ref, bs, round_mode=1, **p)
        assert np.array_equal(fused["recon"], golden[f"pf_{n}_final_r"])
        assert np.array_equal(fused["planes"], golden[f"pf_{n}_planes_r"])


@pytest.mark.parametrize("qf", [10, 50, 99])
def test_stills_quality_sweep(orc, golden, golden_meta, qf):
    """DCTCompression/dct.py:169-208 at QF 10/50/99 (BASELINE config 4's DCT part)."""
    img = golden["still_img"]
    Q = orc.qtables(float(qf))
    planes = orc.compress(img, Q)
    assert np.array_equal(planes, golden[f"still_q{qf}_planes"])
    pr = orc.compress(img, Q, round_mode=1)
    assert np.array_equal(pr, np.round(planes))
    assert np.array_equal(orc.decompress(pr, Q), golden[f"still_q{qf}_dec"])
    sparsity = 1.0 - np.count_nonzero(pr) / pr.size
    assert sparsity == golden_meta[f"still_q{qf}_sparsity"]


def test_full_clip_pins(golden_meta):
    """Recorded in the build container: the oracle reproduced the reference's MVs on all 114
    P-frames of videos/traffic_cut.mp4 (sha pin of SURVEY 4) and whole P-frames bit-exactly."""
    assert golden_meta["traffic_full_mv_sha16"] == "ddae5b429d9f442e"
    assert golden_meta["traffic_full_oracle_mv_sha16"] == golden_meta["traffic_full_mv_sha16"]
    for n in (1, 35):
        assert all(v == 0 for v in golden_meta[f"traffic_frame{n}_mismatch"].values())


def test_symmetric_search_properties(orc):
    """Generalised +/-R step-1 mode (no literal oracle): it must (i) contain the reference's
    step-1 interval as a subset and agree with it whenever the winner lies inside, (ii) find a
    planted shift exactly under SAD."""
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    ref = base[16:80, 16:112].copy()
    cur = base[16 + 5:80 + 5, 16 - 7:112 - 7].copy()     # cur(y,x) = ref(y+5, x-7)
    mv, cost, flags = orc.me(cur, ref, 16, metric=orc.METRIC_SAD, static_thr=-1,
                             **orc.symmetric_search_params(16))
    H, W = 64, 96
    coords = orc.block_coords(H, W, 16)
    inner = (coords[:, 0] - 7 >= 0) & (coords[:, 1] + 5 + 16 <= H)
    assert np.all(mv[inner] == [-7, 5]) and np.all(cost[inner] == 0)
    # subset agreement with the literal interval, wrap8 metric
    lit = orc.reference_search_params(16, R=16, step=1)      # dy in [-16,-1]
    mv_l, cost_l, _ = orc.me(cur, ref, 16, static_thr=-1, **lit)
    mv_s, cost_s, _ = orc.me(cur, ref, 16, static_thr=-1, **orc.symmetric_search_params(16))
    assert np.all(cost_s <= cost_l)
    same = np.all((mv_s >= lit["lo"]) & (mv_s <= lit["hi"]), axis=1)
    # where the symmetric winner is inside the literal interval and not clipped by slack
    for k in np.nonzero(same)[0]:
        if cost_s[k] == cost_l[k]:
            assert tuple(mv_s[k]) == tuple(mv_l[k])


def test_traffic_cut_pins(orc):
    """BASELINE configs[0] on the real clip: the oracle against sha256 pins taken from the UNMODIFIED reference
    (tests/golden/make_golden_traffic.py) -- all 114 x 3600 motion vectors, static counts, and the float64 planes and
    decoded frames of P-frames 1 and 35.  The clip is a data fixture copied from the reference's videos/."""
    import hashlib
    import json
    import os
    cv2 = pytest.importorskip("cv2")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, "golden_traffic_meta.json")) as f:
        meta = json.load(f)
    cap = cv2.VideoCapture(os.path.join(here, meta["file"]))
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    cap.release()

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    if len(frames) != meta["frames"] or sha(np.stack(frames)) != meta["frames_sha16"]:
        pytest.skip("this host's video decoder yields different pixels than the one the pins were taken with")
    prm = orc.reference_search_params(8)
    mvs = [orc.me(f, frames[(n // 4) * 4], 8, **prm)[0] for n, f in enumerate(frames) if n % 4]
    mv = np.stack(mvs).astype(np.int32)
    assert sha(mv) == meta["mv_sha16"]
    statics = [int(((m[:, 0] == 0) & (m[:, 1] == 0)).sum()) for m in mv]
    assert [min(statics), max(statics)] == meta["static_minmaxmean"][:2]
    for n in (1, 35):
        o = orc.encode_p(frames[n], frames[(n // 4) * 4], 8, **prm)
        assert sha(o["planes"]) == meta[f"frame{n}_planes_sha16"]
        assert sha(o["recon"]) == meta[f"frame{n}_final_sha16"]
