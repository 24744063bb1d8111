"""Parity of the CUDA path (through the C ABI) against the reference's golden vectors and the
CPU oracle.  Integer outputs (MVs, costs, flags, pixels, quantised indices) and the float64
coefficient planes must be bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vcs():
    import vcs_h264_b200 as v
    v.runtime.get_context()          # fails loudly without a GPU / the built extension
    return v


def _gpu_me(vcs, cur, ref, bs, lo, hi, step, slack, metric=0, static_thr=2000, kernel=0):
    c = vcs._capi
    H, W = cur.shape[:2]
    p = c.me_reference_params(H, W, bs)
    p.lo, p.hi, p.step, p.slack, p.metric, p.static_thr, p.kernel = lo, hi, step, slack, metric, static_thr, kernel
    N = c.num_blocks(H, W, bs)
    mv = np.empty((N, 2), np.int16); cost = np.empty(N, np.uint32); flags = np.empty(N, np.uint8)
    ctx = vcs.runtime.get_context()
    cur = np.ascontiguousarray(cur); ref = np.ascontiguousarray(ref)
    ctx.call("vcs_me_search_host", p, cur.ctypes.data, ref.ctypes.data, mv.ctypes.data,
             cost.ctypes.data, flags.ctypes.data)
    return mv.astype(np.int32), cost, flags


def test_me_golden_cases(vcs, orc, golden, golden_meta):
    """Every golden ME case of the unmodified reference (motion.py:20-36), both kernels."""
    for case in golden_meta["me_cases"]:
        n, bs = case["name"], case["bs"]
        cur, ref = golden[f"me_{n}_cur"], golden[f"me_{n}_ref"]
        prm = orc.reference_search_params(bs, R=case["R"], step=1 if case["step1"] else None)
        for kernel in (vcs.ME_GENERIC, vcs.ME_AUTO):
            mv, cost, flags = _gpu_me(vcs, cur, ref, bs, kernel=kernel, **prm)
            assert np.array_equal(mv, golden[f"me_{n}_mv"]), (n, kernel)
            assert np.array_equal(flags & 1, golden[f"me_{n}_static"]), (n, kernel)
            gc = golden[f"me_{n}_cost"]
            ok = gc >= 0
            assert np.array_equal(cost[ok].astype(np.int64), gc[ok]), (n, kernel)
            assert np.all(flags[~ok] == 2), (n, kernel)
            omv, ocost, oflags = orc.me(cur, ref, bs, **prm)
            assert np.array_equal(cost, ocost) and np.array_equal(flags, oflags)


@pytest.mark.parametrize("bs,R", [(16, 16), (16, 8), (8, 8), (8, 16), (16, 32), (4, 4)])
@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("thr", [2000, -1])
def test_full_search_vs_oracle(vcs, orc, bs, R, metric, thr):
    """Symmetric +/-R step-1 search (BASELINE configs 2/3/5) on seeded frames, incl. frame-edge
    clipping, vs the oracle restatement; MVs, costs and flags bit-exact."""
    from vcs_h264_b200 import synth
    H, W = (112, 176) if bs >= 8 else (40, 56)
    clip = synth.clip(3, H, W, seed=bs * 100 + R, noise=2, margin=48)
    rng = np.random.default_rng(R)
    cur, ref = clip[2].copy(), clip[0]
    cur[:bs * 2, :bs * 3] = ref[:bs * 2, :bs * 3]            # some static macroblocks
    cur[-bs:, -bs * 2:] = rng.integers(0, 256, (bs, bs * 2, 3), dtype=np.uint8)
    prm = orc.symmetric_search_params(R)
    omv, ocost, oflags = orc.me(cur, ref, bs, metric=metric, static_thr=thr, **prm)
    for kernel in (vcs.ME_GENERIC, vcs.ME_AUTO):
        mv, cost, flags = _gpu_me(vcs, cur, ref, bs, metric=metric, static_thr=thr, kernel=kernel, **prm)
        assert np.array_equal(mv, omv), kernel
        assert np.array_equal(cost, ocost), kernel
        assert np.array_equal(flags, oflags), kernel


def test_ties_first_minimum_wins(vcs, orc):
    """Flat and periodic frames: many exact ties; scan order rows-outer/cols-inner, strict '<'
    (motion.py:133-152) must pick the same candidate as the oracle."""
    yy, xx = np.mgrid[0:96, 0:160]
    per = (((xx % 4) * 50 + (yy % 2) * 30) % 256).astype(np.uint8)
    ref = np.stack([per, per, per], -1)
    cur = np.roll(ref, (2, 4), (0, 1)).copy()
    flat_r = np.full((96, 160, 3), 7, np.uint8); flat_c = np.full((96, 160, 3), 200, np.uint8)
    for c_, r_ in ((cur, ref), (flat_c, flat_r), (flat_r, flat_c)):
        for metric in (0, 1):
            for bs, R in ((16, 16), (8, 8)):
                prm = orc.symmetric_search_params(R)
                omv, ocost, ofl = orc.me(c_, r_, bs, metric=metric, static_thr=-1, **prm)
                for kernel in (vcs.ME_GENERIC, vcs.ME_AUTO):
                    mv, cost, fl = _gpu_me(vcs, c_, r_, bs, metric=metric, static_thr=-1, kernel=kernel, **prm)
                    assert np.array_equal(mv, omv) and np.array_equal(cost, ocost)


def test_p_frame_pipeline_golden(vcs, golden, golden_meta):
    """encoder.py:49-70 + decoder.py:52-69 through the drop-in classes, stage by stage, against
    the reference's own outputs (float64 planes bit-exact)."""
    for case in golden_meta["pf_cases"]:
        n, bs = case["name"], case["bs"]
        cur, ref = golden[f"pf_{n}_cur"], golden[f"pf_{n}_ref"]
        H, W = cur.shape[:2]
        mp = vcs.MotionProcessor(bs, [H, W])
        dc = vcs.DCTCompressor(8)
        mvs, coords = mp.process_motion_prediction(cur, ref)
        assert isinstance(mvs, list) and isinstance(mvs[0][0], int)
        assert np.array_equal(np.array(mvs), golden[f"pf_{n}_mv"])
        pred = mp.reconstruct_from_motion_vectors(mvs, ref, coords)
        assert np.array_equal(pred, golden[f"pf_{n}_pred"])
        resid = mp.get_residuals(input_frame=cur, reconstructed=pred)
        assert np.array_equal(resid, golden[f"pf_{n}_resid"])
        planes = dc.compress(resid)
        assert len(planes) == 3 and planes[0].dtype == np.float64
        assert np.array_equal(np.stack(planes), golden[f"pf_{n}_planes"])
        dec = dc.decompress(compressed=planes, imshape=pred.shape)
        assert np.array_equal(dec, golden[f"pf_{n}_dec"])
        assert np.array_equal(mp._add(pred, dec), golden[f"pf_{n}_final"])
        planes_r = dc.compress(resid, rounded=True)
        assert np.array_equal(np.stack(planes_r), golden[f"pf_{n}_planes_r"])
        assert np.array_equal(dc.decompress(planes_r, pred.shape, pred=pred), golden[f"pf_{n}_final_r"])
        idx = dc.compress_indices(resid)
        assert np.array_equal(idx.astype(np.float64), golden[f"pf_{n}_planes_r"])
        assert np.array_equal(dc.decompress(list(idx), pred.shape), golden[f"pf_{n}_dec_r"])


@pytest.mark.parametrize("qf", [10, 50, 99])
def test_stills_quality_sweep(vcs, golden, golden_meta, qf):
    """DCTCompression/dct.py:169-208 at QF 10/50/99."""
    from vcs_h264_b200.DCTcompressor import quality_tables
    img = golden["still_img"]
    dc = vcs.DCTCompressor(8)
    dc.Q = quality_tables(qf)
    planes = dc.compress(img)
    assert np.array_equal(np.stack(planes), golden[f"still_q{qf}_planes"])
    pr = dc.compress(img, rounded=True)
    assert np.array_equal(dc.decompress(pr, img.shape), golden[f"still_q{qf}_dec"])
    sparsity = 1.0 - np.count_nonzero(np.stack(pr)) / np.stack(pr).size
    assert sparsity == golden_meta[f"still_q{qf}_sparsity"]


def test_encoder_decoder_dropin(vcs, orc):
    """main.py's loop on a small synthetic clip: Encoder/Decoder drop-ins vs the oracle."""
    from vcs_h264_b200 import synth
    clip = synth.clip(6, 64, 96, seed=5, margin=32)
    enc = vcs.Encoder(pattern=["I", "P", "P", "P"], shape=[64, 96], block_size=8, with_DCT=True)
    for n, f in enumerate(clip):
        enc.encode_frame(f, n)
    assert [f.t for f in enc.encoded_frames] == ["I", "P", "P", "P", "I", "P"]
    dec = vcs.Decoder(encoded_frames=enc.encoded_frames, fps=25.0, shape=[64, 96],
                      ref_frames=enc.ref_frames, block_size=8, with_DCT=True)
    frames = dec.decode_frames(with_residuals=True)
    prm = orc.reference_search_params(8)
    for n, f in enumerate(clip):
        if n % 4 == 0:
            assert np.array_equal(frames[n], f)
            continue
        o = orc.encode_p(f, clip[(n // 4) * 4], 8, **prm)
        fr = enc.encoded_frames[n]
        assert np.array_equal(np.array(fr.mv), o["mv"])
        assert np.array_equal(np.stack(fr.r), o["planes"])
        assert np.array_equal(frames[n], o["recon"])
        assert fr.ref_i == n // 4 and fr.i == n


@pytest.mark.parametrize("coef_mode", [0, 1, 2, 3])
def test_clip_api_host_and_device(vcs, orc, coef_mode):
    """Whole-clip C-ABI calls (host buffers and device tensors) vs per-frame oracle."""
    import torch
    from vcs_h264_b200 import synth
    T, H, W, bs, R = 10, 64, 96, 16, 8
    clip = synth.clip(T, H, W, seed=77, margin=32)
    clip[5, :32, :48] = clip[4, :32, :48]                     # static blocks in one P-frame
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=4,
                         coef_mode=coef_mode)
    out = ce.encode_host(clip, want_coef=True, want_recon=True)
    dev_in = torch.from_numpy(clip).cuda()
    dout = ce.alloc_device_outputs(T)
    ce.encode_device(dev_in, dout)
    torch.cuda.synchronize()
    prm = orc.symmetric_search_params(R)
    for p, t in enumerate(ce.p_frame_indices(T)):
        o = orc.encode_p(clip[t], clip[(t // 4) * 4], bs, round_mode=int(coef_mode != 0), **prm)
        for res in (out, {k: v.cpu() for k, v in dout.items()}):
            assert np.array_equal(np.asarray(res["mv"][p]).astype(np.int32), o["mv"])
            assert np.array_equal(np.asarray(res["cost"][p]).view(np.uint32), o["cost"])
            assert np.array_equal(np.asarray(res["flags"][p]), o["flags"])
            assert np.array_equal(np.asarray(res["coef"][p]).astype(np.float64), o["planes"])
            assert np.array_equal(np.asarray(res["recon"][p]), o["recon"])


def test_reference_search_clip(vcs, orc):
    """ClipEncoder(search='reference') == MotionProcessor defaults on every P-frame."""
    from vcs_h264_b200 import synth
    T, H, W = 7, 72, 104
    clip = synth.clip(T, H, W, seed=9, margin=32)
    for bs in (8, 16):
        ce = vcs.ClipEncoder([H, W], block_size=bs, search="reference", gop_len=4, coef_mode=0)
        out = ce.encode_host(clip, want_coef=True, want_recon=True)
        prm = orc.reference_search_params(bs)
        for p, t in enumerate(ce.p_frame_indices(T)):
            o = orc.encode_p(clip[t], clip[(t // 4) * 4], bs, **prm)
            assert np.array_equal(np.asarray(out["mv"][p]).astype(np.int32), o["mv"])
            assert np.array_equal(np.asarray(out["coef"][p]), o["planes"])
            assert np.array_equal(np.asarray(out["recon"][p]), o["recon"])


def test_full_size_properties_1080p(vcs, orc):
    """BASELINE config 2 geometry (1080p, bs 16, +/-16): size-independent properties.
    (i) planted global shift is found exactly under SAD with zero cost away from the borders;
    (ii) un-rounded DCT->IDCT is the identity up to the truncating store and the lossy 8-bit
         YCrCb round trip: |recon - cur| stays within a few grey levels;
    (iii) a strip of the frame agrees bit-exactly with the oracle."""
    import torch
    from vcs_h264_b200 import synth
    H, W, bs, R = 1080, 1920, 16, 16
    base = synth.texture(H, W, seed=3, margin=32)
    ref = np.ascontiguousarray(base[32:32 + H, 32:32 + W])
    cur = np.ascontiguousarray(base[32 + 5:32 + 5 + H, 32 - 9:32 - 9 + W])   # cur(y,x)=ref(y+5,x-9)
    clip = np.stack([ref, cur])
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=2,
                         metric=vcs.METRIC_SAD, static_thr=-1, coef_mode=0)
    out = ce.encode_host(clip, want_coef=True, want_recon=True)
    mv = np.asarray(out["mv"][0]).astype(np.int32).reshape(H // bs, W // bs, 2)
    cost = np.asarray(out["cost"][0]).view(np.uint32).reshape(H // bs, W // bs)
    inner = (slice(0, (H - 5 - bs) // bs), slice(1, None))
    assert np.all(mv[inner] == [-9, 5]) and np.all(cost[inner] == 0)
    recon = np.asarray(out["recon"][0]).astype(np.int16)
    d = np.abs(((recon - cur.astype(np.int16) + 128) % 256) - 128)
    assert d.max() <= 6 and d.mean() < 1.5
    # (iii) oracle on a strip: rows 0..95 see candidates only inside rows 0..127
    strip = 96
    ce2 = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=2, coef_mode=2)
    rng = np.random.default_rng(0)
    cur2 = np.clip(cur.astype(np.int16) + rng.integers(-2, 3, cur.shape), 0, 255).astype(np.uint8)
    out2 = ce2.encode_host(np.stack([ref, cur2]), want_coef=True, want_recon=True)
    o = orc.encode_p(cur2[:strip + 32], ref[:strip + 32], bs, round_mode=1,
                     **orc.symmetric_search_params(R))
    nrow = strip // bs
    nbx = W // bs
    assert np.array_equal(np.asarray(out2["mv"][0]).astype(np.int32)[:nrow * nbx], o["mv"][:nrow * nbx])
    assert np.array_equal(np.asarray(out2["cost"][0]).view(np.uint32)[:nrow * nbx], o["cost"][:nrow * nbx])
    assert np.array_equal(np.asarray(out2["coef"][0])[:, :strip].astype(np.float64), o["planes"][:, :strip])
    assert np.array_equal(np.asarray(out2["recon"][0])[:strip], o["recon"][:strip])


def test_shards_equal_single(vcs):
    """N GOP shards == 1 shard, bit for bit."""
    from vcs_h264_b200 import sharding, synth
    T, H, W = 18, 64, 96
    clip = synth.clip(T, H, W, seed=21, margin=32)
    ce = vcs.ClipEncoder([H, W], block_size=16, search="full", search_range=8, gop_len=4)
    whole = ce.encode_host(clip, want_coef=True, want_recon=True)
    for world in (2, 3, 4):
        parts = {k: [] for k in whole}
        for r in range(world):
            t0, t1 = sharding.frame_range(T, 4, r, world)
            if t1 > t0:
                o = ce.encode_host(np.ascontiguousarray(clip[t0:t1]), want_coef=True, want_recon=True)
                for k in whole:
                    parts[k].append(np.asarray(o[k]))
        for k in whole:
            assert np.array_equal(np.concatenate(parts[k], 0), np.asarray(whole[k])), (k, world)


def test_error_behaviour(vcs):
    c = vcs._capi
    ctx = vcs.runtime.get_context()
    with pytest.raises(ValueError):                      # reference: broadcast error for bs != 8
        vcs.DCTCompressor(16).compress(np.zeros((16, 16, 3), np.uint8))
    with pytest.raises(ValueError):                      # sides not multiples of 8
        vcs.DCTCompressor(8).compress(np.zeros((12, 16, 3), np.uint8))
    p = c.me_reference_params(64, 64, 8)
    p.step = 0
    buf = np.zeros((64, 64, 3), np.uint8)
    mv = np.zeros((64, 2), np.int16)
    with pytest.raises(c.VcsError):
        ctx.call("vcs_me_search_host", p, buf.ctypes.data, buf.ctypes.data, mv.ctypes.data, None, None)
    p = c.me_reference_params(64, 64, 8)
    p.kernel = c.ME_TILED                                # step 3: tiled kernel must refuse, not fall back
    with pytest.raises(c.VcsError):
        ctx.call("vcs_me_search_host", p, buf.ctypes.data, buf.ctypes.data, mv.ctypes.data, None, None)
    mp = vcs.MotionProcessor(8, [64, 64])
    with pytest.raises(c.VcsError):                      # MV pointing outside the frame
        mp.reconstruct_from_motion_vectors([[-100, 0]] * 64, buf, mp._block_coords().tolist())


@pytest.mark.parametrize("coef_mode", [0, 2, 3])
def test_clip_decoder_roundtrip(vcs, orc, coef_mode):
    """Decoder side for a whole clip (decoder.py:52-69): MC from the original I-frames + decompress
    + wrap add == the encoder's own reconstruction == the oracle, bit for bit (host and device)."""
    import torch
    from vcs_h264_b200 import synth
    T, H, W, bs, R, g = 9, 64, 96, 16, 8, 4
    clip = synth.clip(T, H, W, seed=41, margin=32)
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=g, coef_mode=coef_mode)
    enc = ce.encode_host(clip, want_coef=True, want_recon=True)
    refs = np.ascontiguousarray(clip[::g])
    cd = vcs.ClipDecoder([H, W], block_size=bs, gop_len=g, coef_mode=coef_mode)
    dec = cd.decode_host(refs, np.asarray(enc["mv"]), np.asarray(enc["coef"]), T)
    assert np.array_equal(dec, np.asarray(enc["recon"]))
    ddev = torch.empty((ce.num_p_frames(T), H, W, 3), dtype=torch.uint8, device="cuda")
    cd.decode_device(torch.from_numpy(refs).cuda(), enc["mv"].cuda(), enc["coef"].cuda(), ddev, T)
    torch.cuda.synchronize()
    assert np.array_equal(ddev.cpu().numpy(), dec)
    prm = orc.symmetric_search_params(R)
    for p, t in enumerate(ce.p_frame_indices(T)):
        o = orc.encode_p(clip[t], clip[(t // g) * g], bs, round_mode=int(coef_mode != 0), **prm)
        assert np.array_equal(dec[p], o["recon"])


@pytest.mark.parametrize("qf", [10, 50, 99])
def test_sparsity_on_device(vcs, golden, golden_meta, qf):
    """The sparsity print of DCTCompression/dct.py:188-191 from a device reduction."""
    import torch
    from vcs_h264_b200.DCTcompressor import quality_tables
    img = golden["still_img"]
    dc = vcs.DCTCompressor(8)
    dc.Q = quality_tables(qf)
    idx = torch.from_numpy(dc.compress_indices(img)).cuda()
    ctx = vcs.runtime.get_context()
    assert vcs.sparsity_device(ctx, idx, vcs.COEF_I16_RINT) == golden_meta[f"still_q{qf}_sparsity"]
    pl = torch.from_numpy(np.stack(dc.compress(img, rounded=True))).cuda()
    assert vcs.sparsity_device(ctx, pl, vcs.COEF_F64_RINT) == golden_meta[f"still_q{qf}_sparsity"]


def test_int8_indices_refused_when_lossy(vcs):
    """VCS_COEF_I8_RINT is only offered when |index| <= 1024/min(Q) fits int8 (min Q >= 9)."""
    from vcs_h264_b200 import synth
    clip = synth.clip(2, 32, 48, seed=1, margin=32)
    for qf, ok in ((50.0, True), (10.0, True), (75.0, False), (99.0, False)):
        ce = vcs.ClipEncoder([32, 48], block_size=16, search="full", search_range=4, gop_len=2, qf=qf,
                             coef_mode=vcs.COEF_I8_RINT)
        if ok:
            out = ce.encode_host(clip)
            ref = vcs.ClipEncoder([32, 48], block_size=16, search="full", search_range=4, gop_len=2, qf=qf,
                                  coef_mode=vcs.COEF_I16_RINT).encode_host(clip)
            assert np.array_equal(np.asarray(out["coef"]).astype(np.int16), np.asarray(ref["coef"]))
        else:
            with pytest.raises(vcs.VcsError):
                ce.encode_host(clip)


def test_main_driver_frame_and_clip_modes_agree(vcs, orc):
    """main.py's loop (main.py:29-50): per-frame drop-in classes and the clip path give the same decoded
    frames as the oracle."""
    from vcs_h264_b200 import main as drv, synth
    clip = synth.clip(7, 64, 96, seed=3, margin=32)
    frames = [clip[t] for t in range(7)]
    _, dec_f = drv.run(frames, block_size=8, mode="frame")
    _, dec_c = drv.run(frames, block_size=8, mode="clip")
    prm = orc.reference_search_params(8)
    for t in range(7):
        want = frames[t] if t % 4 == 0 else orc.encode_p(frames[t], frames[(t // 4) * 4], 8, **prm)["recon"]
        assert np.array_equal(dec_f[t], want) and np.array_equal(dec_c[t], want)


@pytest.mark.gpu
@pytest.mark.parametrize("qf", [50, 25, 10])
def test_quantiser_exact_half_integers(vcs, orc, qf):
    """Flat 8x8 blocks give D[0][0] = 8*(v-128): with even Q entries the quotient is an exact half-integer for
    many grey levels, where the kernel's q0 = D*RN(1/Q) shortcut must hand over to the IEEE quotient and
    np.round's half-to-even (dct.py:179) decides.  Gradients add near-ties on the AC terms."""
    H, W = 128, 256
    rng = np.random.default_rng(qf)
    grey = np.repeat(np.repeat(rng.integers(0, 256, (H // 8, W // 8), dtype=np.uint8), 8, 0), 8, 1)
    img = np.stack([grey, grey, grey], -1)                     # B = G = R -> Y = grey, Cr = Cb = 128
    img[64:] = np.clip(img[64:].astype(int) + (np.arange(W)[None, :, None] % 8) * 2, 0, 255).astype(np.uint8)
    dc = vcs.DCTCompressor(8)
    dc.Q = list(orc.qtables(float(qf)))
    got = np.stack(dc.compress(img, rounded=True))
    want = orc.compress(img, Q=orc.qtables(float(qf)), round_mode=1)
    assert np.array_equal(got, want)
    halves = np.abs(np.abs(orc.compress(img, Q=orc.qtables(float(qf)), round_mode=0)) % 1.0 - 0.5) < 1e-12
    assert halves.sum() > 50                                    # the case under test really occurs
    assert np.array_equal(dc.compress_indices(img), want.astype(np.int16))


@pytest.mark.gpu
def test_full_size_properties_4k_range32(vcs, orc):
    """BASELINE config 3 geometry (2160x3840, bs 16, +/-32 = 4 chunks of the 65x65 offset range per tile):
    (i) a planted shift beyond +/-16 is found exactly (SAD, zero cost away from the borders);
    (ii) top, middle-free and bottom strips agree bit-exactly with the oracle under the reference cost
         with the static test on (wrap8, thr 2000), including the clipped candidate sets at the borders."""
    from vcs_h264_b200 import synth
    H, W, bs, R = 2160, 3840, 16, 32
    base = synth.texture(H, W, seed=11, margin=64)
    ref = np.ascontiguousarray(base[64:64 + H, 64:64 + W])
    cur = np.ascontiguousarray(base[64 - 27:64 - 27 + H, 64 + 30:64 + 30 + W])      # cur(y,x) = ref(y-27, x+30)
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=2,
                         metric=vcs.METRIC_SAD, static_thr=-1, coef_mode=2)
    out = ce.encode_host(np.stack([ref, cur]), want_coef=False, want_recon=False)
    nby, nbx = H // bs, W // bs
    mv = np.asarray(out["mv"][0]).astype(np.int32).reshape(nby, nbx, 2)
    cost = np.asarray(out["cost"][0]).view(np.uint32).reshape(nby, nbx)
    inner = (slice(2, None), slice(0, nbx - 2))
    assert np.all(mv[inner] == [30, -27]) and np.all(cost[inner] == 0)
    # (ii) strips: the first 3 MB rows only see reference rows < 3*16+32; same for the last 3 by symmetry
    rng = np.random.default_rng(5)
    cur2 = np.clip(cur.astype(np.int16) + rng.integers(-2, 3, cur.shape), 0, 255).astype(np.uint8)
    cur2[:16, :64] = ref[:16, :64]                                                   # a few static blocks
    ce2 = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=2, coef_mode=2)
    out2 = ce2.encode_host(np.stack([ref, cur2]), want_coef=False, want_recon=False)
    mv2 = np.asarray(out2["mv"][0]).astype(np.int32).reshape(nby, nbx, 2)
    cost2 = np.asarray(out2["cost"][0]).view(np.uint32).reshape(nby, nbx)
    fl2 = np.asarray(out2["flags"][0]).reshape(nby, nbx)
    rows, need = 3, 3 * bs + R
    top = orc.me(cur2[:need], ref[:need], bs, **orc.symmetric_search_params(R))
    bot = orc.me(cur2[H - need:], ref[H - need:], bs, **orc.symmetric_search_params(R))
    nb_strip = need // bs
    for got_rows, o, sl in ((slice(0, rows), top, slice(0, rows)), (slice(nby - rows, nby), bot, slice(nb_strip - rows, nb_strip))):
        omv = o[0].reshape(nb_strip, nbx, 2)[sl]
        ocost = o[1].reshape(nb_strip, nbx)[sl]
        ofl = o[2].reshape(nb_strip, nbx)[sl]
        assert np.array_equal(mv2[got_rows], omv)
        assert np.array_equal(fl2[got_rows], ofl)
        assert np.array_equal(cost2[got_rows][ofl == 0], ocost[ofl == 0])
    assert fl2[0, :4].all()                                                          # the planted static blocks


@pytest.mark.gpu
@pytest.mark.parametrize("sched", ["", "1", "2", "5", "1,3,2", "100"])
@pytest.mark.parametrize("geom", [(64, 96, 16, 8, 4), (40, 56, 8, 8, 4), (48, 72, 8, 4, 3), (64, 96, 16, 8, 2)])
def test_host_pipeline_segments_equal_device(vcs, monkeypatch, sched, geom):
    """vcs_encode_clip_host cuts the clip into P-frame segments that may start inside a GOP (FrameAddr.p_off);
    every schedule must give the one-launch device result, for the tiled kernel (W % 16 == 0) and the generic one,
    partial trailing GOPs and 1-P GOPs included."""
    import torch
    from vcs_h264_b200 import synth
    H, W, bs, R, gop = geom
    T = 11
    clip = synth.clip(T, H, W, seed=H + W, margin=32)
    if sched:
        monkeypatch.setenv("VCS_PIPELINE_P", sched)
    else:
        monkeypatch.delenv("VCS_PIPELINE_P", raising=False)
    ce = vcs.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=gop, coef_mode=2)
    out = ce.encode_host(clip, want_coef=True, want_recon=True)
    dout = ce.alloc_device_outputs(T)
    ce.encode_device(torch.from_numpy(clip).cuda(), dout)
    torch.cuda.synchronize()
    for k in ("mv", "cost", "flags", "coef", "recon"):
        assert torch.equal(torch.as_tensor(np.asarray(out[k])), dout[k].cpu()), k
