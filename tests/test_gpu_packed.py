"""Packed coefficient form (include/vcs_b200.h: vcs_encode_clip_host_packed, vcs_pack_coef_dev, vcs_unpack_coef_dev,
vcs_decode_clip_host_packed) against the dense int8 planes: an exact re-coding, checked with the host-side byte
shuffler of the container (vcs_h264_b200.container.expand_packed / compact_dense) and by decoding both forms."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _planes(rng, nP, H, W):
    coef = rng.integers(-128, 128, (nP, 3, H, W)).astype(np.int8)
    small = rng.integers(-8, 8, coef.shape).astype(np.int8)
    pick = rng.random(coef.shape) < 0.85
    coef[pick] = small[pick]
    coef[rng.random(coef.shape) < 0.57] = 0
    coef[0, 0, :8, :8] = 0
    coef[-1, 2, -8:, -8:] = -3
    coef[-1, 1, -8:, :8] = 99                              # a block of 64 escapes
    coef[0, 1, 0, :5] = [-8, 7, -9, 8, 1]
    if nP > 1:
        coef[1] = 0                                        # a whole frame of empty blocks
    return coef


@pytest.mark.parametrize("geom", [(1, 8, 8), (2, 40, 72), (3, 64, 8 * 33), (2, 536, 960)])
def test_pack_unpack_device_round_trip(geom):
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import _capi, container
    nP, H, W = geom
    ctx = _capi.Context(0)
    coef = _planes(np.random.default_rng(H * W), nP, H, W)
    d = torch.from_numpy(coef).cuda()
    bitmap = torch.empty((nP, 3, H // 8, W // 8), dtype=torch.int64, device="cuda")
    row_count = torch.empty((nP, 3, H // 8, 2), dtype=torch.int32, device="cuda")
    nibbles = torch.full((coef.size // 2 + bitmap.numel() + 64,), 77, dtype=torch.uint8, device="cuda")
    escapes = torch.full((coef.size + 64,), 77, dtype=torch.int8, device="cuda")
    n = (C.c_uint64 * 2)(0, 0)
    ctx.call("vcs_pack_coef_dev", H, W, nP, d.data_ptr(), bitmap.data_ptr(), row_count.data_ptr(), nibbles.data_ptr(),
             escapes.data_ptr(), n)
    wb, wr, wn, we = container.compact_dense(coef)
    assert (n[0], n[1]) == (wn.size, we.size)
    assert np.array_equal(bitmap.cpu().numpy().view(np.uint64), wb)
    assert np.array_equal(row_count.cpu().numpy().view(np.uint32), wr)
    assert np.array_equal(nibbles[:n[0]].cpu().numpy(), wn)
    assert np.array_equal(escapes[:n[1]].cpu().numpy(), we)
    assert bool((nibbles[n[0]:] == 77).all()) and bool((escapes[n[1]:] == 77).all())      # nothing written past the streams
    back = torch.full_like(d, 55)
    ctx.call("vcs_unpack_coef_dev", H, W, nP, bitmap.data_ptr(), row_count.data_ptr(), nibbles.data_ptr(), n[0],
             escapes.data_ptr(), n[1], back.data_ptr())
    ctx.synchronize()
    assert torch.equal(back, d)
    if n[0] > 10:                                          # truncated streams are refused, not read past their end
        ctx.call("vcs_unpack_coef_dev", H, W, nP, bitmap.data_ptr(), row_count.data_ptr(), nibbles.data_ptr(), n[0] - 10,
                 escapes.data_ptr(), n[1], back.data_ptr())
        with pytest.raises(v.VcsError):
            ctx.synchronize()
    if n[1] > 3:
        ctx.call("vcs_unpack_coef_dev", H, W, nP, bitmap.data_ptr(), row_count.data_ptr(), nibbles.data_ptr(), n[0],
                 escapes.data_ptr(), n[1] - 3, back.data_ptr())
        with pytest.raises(v.VcsError):
            ctx.synchronize()
    ctx.close()


@pytest.mark.parametrize("T,H,W,bs,R,qf", [(13, 96, 160, 16, 16, 50.0), (60, 1080 // 4 // 8 * 8, 1920 // 4, 16, 16, 50.0),
                                            (5, 72, 104, 8, 8, 50.0), (9, 64, 96, 16, 8, 20.0)])
def test_clip_packed_equals_dense(T, H, W, bs, R, qf):
    """The pipelined host path with the packed sink returns the same vectors and, expanded, the same indices as the dense
    path; both decoders reconstruct the same frames from it."""
    import vcs_h264_b200 as v
    from vcs_h264_b200 import container, synth
    clip = synth.clip(T, H, W, seed=T * 7 + H, margin=64)
    ce = v.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=4, qf=qf, coef_mode=v.COEF_I8_RINT)
    dense = ce.encode_host(clip, want_coef=True, want_recon=True)
    for want_recon in (False, True):          # forward-only: the DCT stage emits bitmaps and counts; with recon: count kernel
        pk = ce.encode_host_packed(clip, want_recon=want_recon)
        coef = np.asarray(dense["coef"])
        assert np.array_equal(np.asarray(pk["mv"]), np.asarray(dense["mv"]))
        assert np.array_equal(np.asarray(pk["flags"]), np.asarray(dense["flags"]))
        if want_recon:
            assert np.array_equal(np.asarray(pk["recon"]), np.asarray(dense["recon"]))
        wb, wr, wn, we = container.compact_dense(coef)
        assert pk["lengths"] == (wn.size, we.size)
        assert np.array_equal(np.asarray(pk["bitmap"]).view(np.uint64), wb)
        assert np.array_equal(np.asarray(pk["row_count"]).view(np.uint32), wr)
        assert np.array_equal(np.asarray(pk["nibbles"])[:wn.size], wn)
        assert np.array_equal(np.asarray(pk["escapes"])[:we.size], we)
    cd = v.ClipDecoder([H, W], block_size=bs, gop_len=4, qf=qf, coef_mode=v.COEF_I8_RINT)
    rec = cd.decode_host_packed(clip[::4], pk["mv"], pk["bitmap"], pk["row_count"], pk["nibbles"], pk["escapes"], pk["lengths"], T)
    assert np.array_equal(rec, np.asarray(dense["recon"]))
    # through the version-2 container
    blob = container.pack_packed(clip[::4], pk["mv"], pk["bitmap"], pk["row_count"], pk["nibbles"], pk["escapes"], pk["lengths"],
                                 T=T, block_size=bs, gop_len=4, qf=qf)
    u = container.unpack(blob)
    rec2 = cd.decode_host_packed(u["i_frames"], u["mv"], u["bitmap"], u["row_count"], u["nibbles"], u["escapes"], u["lengths"], T)
    assert np.array_equal(rec2, rec)
    assert v.ClipEncoder.packed_bytes(pk) < coef.size


def test_small_escape_buffer_is_refused():
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W = 5, 64, 96
    clip = synth.clip(T, H, W, seed=3, margin=48)
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=8, gop_len=4, qf=50.0, coef_mode=v.COEF_I8_RINT)
    out = ce.alloc_host_packed(T, pinned=False, escape_fraction=0.0)       # 64 bytes only
    with pytest.raises(v.VcsError, match="too small"):
        ce.encode_host_packed(clip, out)


def test_damaged_packed_clip_is_refused():
    """Decoding a packed clip whose row counts do not add up to the stream lengths is refused on the host; one whose
    bitmaps claim more indices than its streams hold never reads past their end and fails on the device check."""
    import vcs_h264_b200 as v
    from vcs_h264_b200 import synth
    T, H, W = 5, 64, 96
    clip = synth.clip(T, H, W, seed=11, margin=48)
    ce = v.ClipEncoder([H, W], block_size=16, search="full", search_range=8, gop_len=4, qf=50.0, coef_mode=v.COEF_I8_RINT)
    pk = ce.encode_host_packed(clip, want_recon=True)
    cd = v.ClipDecoder([H, W], block_size=16, gop_len=4, qf=50.0, coef_mode=v.COEF_I8_RINT)
    args = lambda **kw: [kw.get(k, pk[k]) for k in ("mv", "bitmap", "row_count", "nibbles", "escapes", "lengths")]
    good = cd.decode_host_packed(clip[::4], *args(), T)
    assert np.array_equal(good, np.asarray(pk["recon"]))
    rc = np.array(np.asarray(pk["row_count"]), copy=True)
    rc.reshape(-1)[0] += 1
    with pytest.raises(v.VcsError, match="do not add up"):
        cd.decode_host_packed(clip[::4], *args(row_count=rc), T)
    bm = np.array(np.asarray(pk["bitmap"]), copy=True)
    bm.reshape(-1).view(np.uint8)[-8:] = 0xFF                  # the last block now claims 64 indices
    with pytest.raises(v.VcsError):
        cd.decode_host_packed(clip[::4], *args(bitmap=bm), T)
    again = cd.decode_host_packed(clip[::4], *args(), T)          # the context is usable afterwards
    assert np.array_equal(again, good)
