"""Packed container of an encoded clip: byte layout round trip, header validation, and the rebuilt
reference-style Frame objects.  The GPU test feeds a container to both decoders."""
import numpy as np
import pytest

from vcs_h264_b200 import _capi, container


def _random_clip(rng, T=7, H=32, W=48, bs=8, gop=3, cm=_capi.COEF_I16_RINT):
    n_i, n_p, N = (T + gop - 1) // gop, _capi.num_p_frames(T, gop), (H // bs) * (W // bs)
    i_frames = rng.integers(0, 256, (n_i, H, W, 3), dtype=np.uint8)
    # in-frame vectors only: unpack() refuses a vector that would move its macroblock out of the frame
    k = np.arange(N)
    x, y = (k % (W // bs)) * bs, (k // (W // bs)) * bs
    tx, ty = rng.integers(0, W - bs + 1, (n_p, N)), rng.integers(0, H - bs + 1, (n_p, N))
    mv = np.stack([tx - x, ty - y], -1).astype(np.int16)
    dt = container.COEF_DTYPES[cm]
    coef = rng.integers(-100, 101, (n_p, 3, H, W)).astype(dt) if dt != np.float64 else rng.standard_normal((n_p, 3, H, W))
    return dict(i_frames=i_frames, mv=mv, coef=coef, T=T, block_size=bs, gop_len=gop, coef_mode=cm)


@pytest.mark.parametrize("cm", [_capi.COEF_F64, _capi.COEF_I16_RINT, _capi.COEF_I8_RINT])
def test_round_trip(cm):
    c = _random_clip(np.random.default_rng(cm), cm=cm)
    blob = container.pack(c["i_frames"], c["mv"], c["coef"], T=c["T"], block_size=c["block_size"], gop_len=c["gop_len"],
                          coef_mode=cm, qf=75.0)
    u = container.unpack(blob)
    assert (u["T"], u["H"], u["W"], u["block_size"], u["gop_len"], u["coef_mode"], u["qf"]) == (7, 32, 48, 8, 3, cm, 75.0)
    for k in ("i_frames", "mv", "coef"):
        assert u[k].dtype == c[k].dtype and np.array_equal(u[k], c[k])
    assert np.array_equal(u["Q"], _capi.q_tables(75.0))
    itemsize = np.dtype(container.COEF_DTYPES[cm]).itemsize
    assert len(blob) == 64 + 1536 + c["i_frames"].size + c["mv"].size * 2 + c["coef"].size * itemsize


def test_rejects_damaged_containers():
    c = _random_clip(np.random.default_rng(1))
    blob = container.pack(c["i_frames"], c["mv"], c["coef"], T=c["T"], block_size=8, gop_len=3, coef_mode=c["coef_mode"])
    with pytest.raises(ValueError):
        container.unpack(blob[:-1])
    with pytest.raises(ValueError):
        container.unpack(b"X" + blob[1:])
    with pytest.raises(ValueError):
        container.unpack(blob[:40])
    with pytest.raises(ValueError):
        container.pack(c["i_frames"][:1], c["mv"], c["coef"], T=c["T"], block_size=8, gop_len=3, coef_mode=c["coef_mode"])
    bad = c["mv"].copy()
    bad[1, 3, 0] = 1000                      # a vector that leaves the frame (corrupt or foreign input)
    with pytest.raises(ValueError, match="outside the frame"):
        container.unpack(container.pack(c["i_frames"], bad, c["coef"], T=c["T"], block_size=8, gop_len=3,
                                        coef_mode=c["coef_mode"]))


def test_frames_view_matches_the_reference_objects():
    c = _random_clip(np.random.default_rng(2), cm=_capi.COEF_F64)
    u = container.unpack(container.pack(c["i_frames"], c["mv"], c["coef"], T=7, block_size=8, gop_len=3,
                                        coef_mode=_capi.COEF_F64))
    frames, refs = container.to_frames(u)
    assert [f.t for f in frames] == ["I", "P", "P", "I", "P", "P", "I"]
    assert [f.i for f in frames] == list(range(7)) and [f.ref_i for f in frames] == [0, 0, 0, 0, 1, 1, 0]
    assert frames[0].mv is None and frames[0].r is None and len(refs) == 3
    p1 = frames[1]
    assert isinstance(p1.mv, list) and isinstance(p1.mv[0], list) and isinstance(p1.mv[0][0], int)   # [dx, dy] ints
    assert p1.c[:3] == [[0, 0], [8, 0], [16, 0]] and len(p1.c) == 24                                 # [x, y] raster
    assert len(p1.r) == 3 and p1.r[0].dtype == np.float64 and np.array_equal(p1.r[2], c["coef"][0, 2])
    # and back
    blob2 = container.from_frames(frames, refs, block_size=8, gop_len=3, coef_mode=_capi.COEF_F64)
    u2 = container.unpack(blob2)
    assert np.array_equal(u2["mv"], c["mv"]) and np.array_equal(u2["coef"], c["coef"]) and np.array_equal(u2["i_frames"], c["i_frames"])


@pytest.mark.gpu
def test_container_feeds_both_decoders():
    import contextlib
    import io
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W, bs, gop = 7, 64, 96, 8, 3
    clip = synth.clip(T, H, W, seed=9, margin=32)
    ce = v.ClipEncoder([H, W], block_size=bs, search="reference", gop_len=gop, coef_mode=v.COEF_F64)
    out = ce.encode_host(clip, want_coef=True, want_recon=True)
    blob = container.pack(clip[::gop], np.asarray(out["mv"]), np.asarray(out["coef"]), T=T, block_size=bs, gop_len=gop,
                          coef_mode=v.COEF_F64)
    u = container.unpack(blob)
    rec = v.ClipDecoder([H, W], block_size=bs, gop_len=gop, coef_mode=v.COEF_F64).decode_host(u["i_frames"], u["mv"], u["coef"], T)
    assert np.array_equal(rec, np.asarray(out["recon"]))
    frames, refs = container.to_frames(u)
    with contextlib.redirect_stdout(io.StringIO()):
        dec = v.Decoder(encoded_frames=frames, fps=25.0, shape=[H, W], ref_frames=refs, block_size=bs, with_DCT=True,
                        dct_block_size=8)
        decoded = dec.decode_frames(with_residuals=True)
    p = 0
    for t in range(T):
        if t % gop == 0:
            assert np.array_equal(decoded[t], clip[t])
        else:
            assert np.array_equal(decoded[t], rec[p])
            p += 1


def test_packed_form_round_trip():
    """Version-2 container: bitmap + 4-bit codes + escapes is an exact re-coding of the dense int8 indices."""
    rng = np.random.default_rng(7)
    c = _random_clip(rng, T=5, H=40, W=72, bs=8, gop=3, cm=_capi.COEF_I8_RINT)
    coef = c["coef"].copy()
    small = rng.integers(-8, 8, coef.shape).astype(np.int8)
    pick = rng.random(coef.shape) < 0.8                    # most indices are small, like on real residuals
    coef[pick] = small[pick]
    coef[rng.random(coef.shape) < 0.6] = 0                 # the bench clip is 57 % zeros
    coef[0, 0, :8, :8] = 0                                 # an empty block
    coef[1, 2, 8:16, 16:24] = 5                            # a full one
    coef[2, 1, :8, :8] = -8                                # the edges of the code range ...
    coef[2, 1, 8:16, :8] = 7
    coef[2, 1, 16:24, :8] = -9                             # ... and the first escapes on either side
    coef[2, 1, 24:32, :8] = 8
    coef[2, 0, 0, :3] = [-128, 127, 1]                     # an odd number of values in a block: padded to a byte
    bitmap, row_count, nibbles, escapes = container.compact_dense(coef)
    assert bitmap.dtype == np.uint64 and bitmap.shape == (3, 3, 5, 9) and row_count.shape == (3, 3, 5, 2)
    nz = coef != 0
    assert escapes.size == np.count_nonzero(nz & ((coef < -8) | (coef > 7))) == int(row_count[..., 1].sum())
    assert nibbles.size == int(row_count[..., 0].sum())
    assert bitmap[0, 0, 0, 0] == 0 and bitmap[1, 2, 1, 2] == np.uint64(2 ** 64 - 1)
    assert np.array_equal(container.expand_packed(bitmap, row_count, nibbles, escapes, 40, 72), coef)
    blob = container.pack_packed(c["i_frames"], c["mv"], bitmap, row_count, nibbles, escapes, T=5, block_size=8, gop_len=3)
    assert len(blob) == (64 + 1536 + c["i_frames"].size + c["mv"].size * 2 + bitmap.size * 8 + row_count.size * 4
                         + nibbles.size + escapes.size)
    u = container.unpack(blob)
    assert u["version"] == 2 and u["lengths"] == (nibbles.size, escapes.size)
    for k, want in (("bitmap", bitmap), ("row_count", row_count), ("nibbles", nibbles), ("escapes", escapes)):
        assert np.array_equal(u[k], want), k
    frames, refs = container.to_frames(u)
    assert np.array_equal(frames[1].r[0], coef[0, 0].astype(np.float64)) and frames[1].r[0].dtype == np.float64
    with pytest.raises(ValueError):
        container.unpack(blob[:-1])
    with pytest.raises(ValueError):
        container.expand_packed(bitmap, row_count, nibbles[:-1], escapes, 40, 72)
    with pytest.raises(ValueError):
        container.expand_packed(bitmap, row_count, nibbles, escapes[:-1], 40, 72)
