"""Packed coefficient form (include/vcs_b200.h: vcs_encode_clip_host_packed, vcs_pack_coef_dev, vcs_unpack_coef_dev,
vcs_decode_clip_host_packed) against the dense int8 planes: an exact re-coding, checked with the host-side byte
shuffler of the container (vcs_h264_b200.container.expand_packed / compact_dense) and by decoding both forms."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("geom", [(1, 8, 8), (2, 40, 72), (3, 64, 8 * 33), (2, 536, 960)])
def test_pack_unpack_device_round_trip(geom):
    import torch
    import vcs_h264_b200 as v
    from vcs_h264_b200 import _capi, container
    nP, H, W = geom
    ctx = _capi.Context(0)
    rng = np.random.default_rng(H * W)
    coef = rng.integers(-100, 101, (nP, 3, H, W)).astype(np.int8)
    coef[rng.random(coef.shape) < 0.57] = 0
    coef[0, 0, :8, :8] = 0
    coef[-1, 2, -8:, -8:] = -3
    if nP > 1:
        coef[1] = 0                                        # a whole frame of empty blocks
    d = torch.from_numpy(coef).cuda()
    bitmap = torch.empty((nP, 3, H // 8, W // 8), dtype=torch.int64, device="cuda")
    row_count = torch.empty((nP, 3, H // 8), dtype=torch.int32, device="cuda")
    values = torch.full((coef.size + 64,), 77, dtype=torch.int8, device="cuda")
    n = C.c_uint64(0)
    ctx.call("vcs_pack_coef_dev", H, W, nP, d.data_ptr(), bitmap.data_ptr(), row_count.data_ptr(), values.data_ptr(), C.byref(n))
    wb, wr, wv = container.compact_dense(coef)
    assert n.value == wv.size == np.count_nonzero(coef)
    assert np.array_equal(bitmap.cpu().numpy().view(np.uint64), wb)
    assert np.array_equal(row_count.cpu().numpy().view(np.uint32), wr)
    assert np.array_equal(values[:n.value].cpu().numpy(), wv)
    assert bool((values[n.value:] == 77).all())            # nothing written past the stream
    back = torch.full_like(d, 55)
    ctx.call("vcs_unpack_coef_dev", H, W, nP, bitmap.data_ptr(), row_count.data_ptr(), values.data_ptr(), n.value, back.data_ptr())
    ctx.synchronize()
    assert torch.equal(back, d)
    if n.value > 10:                                       # a truncated stream is refused, not read past its end
        ctx.call("vcs_unpack_coef_dev", H, W, nP, bitmap.data_ptr(), row_count.data_ptr(), values.data_ptr(), n.value - 10, back.data_ptr())
        with pytest.raises(v.VcsError):
            ctx.synchronize()
    ctx.close()


@pytest.mark.parametrize("T,H,W,bs,R", [(13, 96, 160, 16, 16), (60, 1080 // 4 // 8 * 8, 1920 // 4, 16, 16), (5, 72, 104, 8, 8)])
def test_clip_packed_equals_dense(T, H, W, bs, R):
    """The pipelined host path with the packed sink returns the same vectors and, expanded, the same indices as the dense
    path; both decoders reconstruct the same frames from it."""
    import vcs_h264_b200 as v
    from vcs_h264_b200 import container, synth
    clip = synth.clip(T, H, W, seed=T * 7 + H, margin=64)
    ce = v.ClipEncoder([H, W], block_size=bs, search="full", search_range=R, gop_len=4, qf=50.0, coef_mode=v.COEF_I8_RINT)
    dense = ce.encode_host(clip, want_coef=True, want_recon=True)
    pk = ce.encode_host_packed(clip, want_recon=True)
    coef = np.asarray(dense["coef"])
    assert np.array_equal(np.asarray(pk["mv"]), np.asarray(dense["mv"]))
    assert np.array_equal(np.asarray(pk["flags"]), np.asarray(dense["flags"]))
    assert np.array_equal(np.asarray(pk["recon"]), np.asarray(dense["recon"]))
    assert pk["nvalues"] == np.count_nonzero(coef)
    vals = np.asarray(pk["values"])[:pk["nvalues"]]
    assert np.array_equal(container.expand_packed(np.asarray(pk["bitmap"]), np.asarray(pk["row_count"]), vals, H, W), coef)
    cd = v.ClipDecoder([H, W], block_size=bs, gop_len=4, qf=50.0, coef_mode=v.COEF_I8_RINT)
    rec = cd.decode_host_packed(clip[::4], pk["mv"], pk["bitmap"], pk["row_count"], pk["values"], pk["nvalues"], T)
    assert np.array_equal(rec, np.asarray(dense["recon"]))
    # through the version-2 container
    blob = container.pack_packed(clip[::4], pk["mv"], pk["bitmap"], pk["row_count"], pk["values"], pk["nvalues"],
                                 T=T, block_size=bs, gop_len=4)
    u = container.unpack(blob)
    rec2 = cd.decode_host_packed(u["i_frames"], u["mv"], u["bitmap"], u["row_count"], u["values"], u["nvalues"], T)
    assert np.array_equal(rec2, rec)
    dense_bytes = coef.size
    packed_bytes = np.asarray(pk["bitmap"]).nbytes + np.asarray(pk["row_count"]).nbytes + pk["nvalues"]
    assert packed_bytes < dense_bytes
