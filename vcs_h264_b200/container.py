"""Packed container of an encoded clip -- the wire form of the reference's list of `Frame` objects.

The reference keeps an encoded clip as Python objects: `Frame(.t .mv .r .c .i .ref_i)` (frame.py:1-8) with the motion
vectors as lists of Python ints, the residual as three float64 planes, and the I-frames kept aside in
`Encoder.ref_frames` (encoder.py:38-43); it has no bitstream.  This module gives that state a byte layout, which is
what crosses NCCL and what a file would hold:

    header  (little endian, 64 bytes): magic "VCSB200\\0", version u32, T u32, H u32, W u32, block_size u32,
            gop_len u32, coef_mode u32, n_p u32, n_blocks u32, qf f64, reserved
    Q       float64[3][8][8]
    I-frames uint8[n_i][H][W][3]                  (frame index t = k * gop_len)
    mv      int16[n_p][n_blocks][2]  ([dx, dy], raster order: MotionProcessor.process_motion_prediction)
    coef    [n_p][3][H][W] in the clip's coefficient format (float64 like DCTCompressor.compress, int16 or int8 indices)

Version 2 stores the int8 indices PACKED (include/vcs_b200.h, vcs_encode_clip_host_packed): instead of `coef` it holds
    bitmap    uint64[n_p][3][H/8][W/8]   occupancy of every 8x8 block (bit 8*i+j = row i, column j)
    row_count uint32[n_p][3][H/8][2]     per block row: bytes of its nibble stream, number of its escapes
    nibbles   uint8[n_nib]               a 4-bit code per non-zero index in block / bit order, low nibble first, blocks padded
                                         to whole bytes; code = v & 15 for v in [-8, 7], 0 = escape
    escapes   int8[n_esc]                the escaped values in the same order
and the header's two reserved fields carry n_nib and n_esc.  `expand_packed()` / `compact_dense()` convert between the two forms on
the host (byte shuffling only, for marshalling into the reference's Frame objects; the decoder's arithmetic stays on the GPU).

`to_frames()` rebuilds the reference's own objects (Frame lists, block coords, ref_frames) from a container, so the
reference's Decoder -- or this package's drop-in Decoder -- can consume it unchanged.  No entropy stage: with the
reference's wrapped residual the indices are dense (43 % non-zero at QF 50 on the bench clip).
"""
from __future__ import annotations

import struct

import numpy as np

from . import _capi
from .frame import Frame

MAGIC = b"VCSB200\0"
VERSION = 1
_HDR = struct.Struct("<8s9IdQI")
COEF_DTYPES = {_capi.COEF_F64: np.float64, _capi.COEF_F64_RINT: np.float64,
               _capi.COEF_I16_RINT: np.int16, _capi.COEF_I8_RINT: np.int8}


def pack(i_frames, mv, coef, *, T, block_size, gop_len, coef_mode, qf=50.0, Q=None) -> bytes:
    """i_frames uint8 [n_i,H,W,3]; mv int16 [n_p,N,2]; coef [n_p,3,H,W] (dtype of coef_mode)."""
    i_frames = np.ascontiguousarray(i_frames, np.uint8)
    n_i, H, W, _ = i_frames.shape
    n_p = _capi.num_p_frames(T, gop_len)
    N = (H // block_size) * (W // block_size)
    mv = np.ascontiguousarray(mv, np.int16)
    coef = np.ascontiguousarray(coef, COEF_DTYPES[coef_mode])
    if n_i != (T + gop_len - 1) // gop_len or mv.shape != (n_p, N, 2) or coef.shape != (n_p, 3, H, W):
        raise ValueError("array shapes do not match T / gop_len / block_size")
    Q = np.ascontiguousarray(_capi.q_tables(qf) if Q is None else Q, np.float64).reshape(3, 8, 8)
    hdr = _HDR.pack(MAGIC, VERSION, T, H, W, block_size, gop_len, coef_mode, n_p, N, float(qf), 0, 0)
    return b"".join([hdr, Q.tobytes(), i_frames.tobytes(), mv.tobytes(), coef.tobytes()])


def pack_packed(i_frames, mv, bitmap, row_count, nibbles, escapes, lengths=None, *, T, block_size, gop_len, qf=50.0, Q=None) -> bytes:
    """Version-2 container: int8 indices in packed form (bitmap uint64 [n_p,3,H/8,W/8], row_count uint32 [n_p,3,H/8,2],
    nibbles uint8 [>= lengths[0]], escapes int8 [>= lengths[1]])."""
    i_frames = np.ascontiguousarray(i_frames, np.uint8)
    n_i, H, W, _ = i_frames.shape
    n_p = _capi.num_p_frames(T, gop_len)
    N = (H // block_size) * (W // block_size)
    mv = np.ascontiguousarray(mv, np.int16)
    bitmap = np.ascontiguousarray(np.asarray(bitmap)).view(np.uint64)
    row_count = np.ascontiguousarray(np.asarray(row_count)).view(np.uint32)
    if row_count.shape != (n_p, 3, H // 8, 2):
        raise ValueError("row_count must be [n_p, 3, H/8, 2]")
    n_nib, n_esc = (int(row_count[..., 0].sum()), int(row_count[..., 1].sum())) if lengths is None else (int(lengths[0]), int(lengths[1]))
    nibbles = np.ascontiguousarray(np.asarray(nibbles).reshape(-1)[:n_nib]).view(np.uint8)
    escapes = np.ascontiguousarray(np.asarray(escapes).reshape(-1)[:n_esc]).view(np.int8)
    if (n_i != (T + gop_len - 1) // gop_len or mv.shape != (n_p, N, 2) or bitmap.shape != (n_p, 3, H // 8, W // 8)
            or nibbles.size != n_nib or escapes.size != n_esc or int(row_count[..., 0].sum()) != n_nib
            or int(row_count[..., 1].sum()) != n_esc or n_esc >= 2 ** 32):
        raise ValueError("array shapes do not match T / gop_len / block_size / stream lengths")
    Q = np.ascontiguousarray(_capi.q_tables(qf) if Q is None else Q, np.float64).reshape(3, 8, 8)
    hdr = _HDR.pack(MAGIC, 2, T, H, W, block_size, gop_len, _capi.COEF_I8_RINT, n_p, N, float(qf), n_nib, n_esc)
    return b"".join([hdr, Q.tobytes(), i_frames.tobytes(), mv.tobytes(), bitmap.tobytes(), row_count.tobytes(),
                     nibbles.tobytes(), escapes.tobytes()])


def expand_packed(bitmap, row_count, nibbles, escapes, H, W):
    """Packed indices -> dense int8 [n_p,3,H,W] (host-side byte shuffling; the GPU twin is vcs_unpack_coef_dev)."""
    bitmap = np.ascontiguousarray(np.asarray(bitmap)).view(np.uint64)
    row_count = np.asarray(row_count).view(np.uint32)
    n_p = bitmap.shape[0]
    bits = np.unpackbits(bitmap.view(np.uint8).reshape(n_p, 3, H // 8, W // 8, 8), axis=-1, bitorder="little").astype(bool)
    bits = bits.reshape(n_p, 3, H // 8, W // 8, 8, 8)      # [.., block row, block column, row in block, column in block]
    nibbles = np.asarray(nibbles).reshape(-1).view(np.uint8)
    escapes = np.asarray(escapes).reshape(-1).view(np.int8)
    n_b = bits.reshape(-1, 64).sum(1).astype(np.int64)     # non-zero indices per block, block order
    nbytes = (n_b + 1) // 2
    if int(nbytes.sum()) != nibbles.size or int(row_count[..., 0].sum()) != nibbles.size or int(row_count[..., 1].sum()) != escapes.size:
        raise ValueError("bitmaps, row counts and streams disagree")
    start = np.cumsum(nbytes) - nbytes                     # first byte of every block's nibble string
    tot = int(n_b.sum())
    blk = np.repeat(np.arange(n_b.size), n_b)              # block of every value, in stream order
    k = np.arange(tot) - np.repeat(np.cumsum(n_b) - n_b, n_b)   # its ordinal inside the block
    code = (nibbles[start[blk] + (k >> 1)] >> (4 * (k & 1)).astype(np.uint8)) & 15
    esc = code == 0
    if int(esc.sum()) != escapes.size:
        raise ValueError("escape codes and escape stream disagree")
    vals = (((code.astype(np.int16) ^ 8) - 8)).astype(np.int8)
    vals[esc] = escapes
    blocks = np.zeros(bits.shape, np.int8)                 # [n_p,3,H/8,W/8,8,8]
    blocks[bits] = vals
    return np.ascontiguousarray(blocks.transpose(0, 1, 2, 4, 3, 5)).reshape(n_p, 3, H, W)


def compact_dense(coef):
    """Dense int8 [n_p,3,H,W] -> (bitmap uint64, row_count uint32 [..,2], nibbles uint8, escapes int8): the inverse of
    expand_packed."""
    coef = np.ascontiguousarray(coef, np.int8)
    n_p, _, H, W = coef.shape
    blocks = coef.reshape(n_p, 3, H // 8, 8, W // 8, 8).transpose(0, 1, 2, 4, 3, 5)
    nz = blocks != 0
    bitmap = np.packbits(nz.reshape(n_p, 3, H // 8, W // 8, 64), axis=-1, bitorder="little").view(np.uint64)[..., 0]
    vals = blocks[nz]                                      # stream order
    small = (vals >= -8) & (vals <= 7)
    code = np.where(small, vals.astype(np.uint8) & 15, 0).astype(np.uint8)
    n_b = nz.reshape(-1, 64).sum(1).astype(np.int64)
    nbytes = (n_b + 1) // 2
    start = np.cumsum(nbytes) - nbytes
    blk = np.repeat(np.arange(n_b.size), n_b)
    k = np.arange(vals.size) - np.repeat(np.cumsum(n_b) - n_b, n_b)
    nibbles = np.zeros(int(nbytes.sum()), np.uint8)
    np.add.at(nibbles, start[blk] + (k >> 1), (code << (4 * (k & 1)).astype(np.uint8)).astype(np.uint8))
    esc_per_block = np.zeros(n_b.size, np.int64)
    np.add.at(esc_per_block, blk, (~small).astype(np.int64))
    shp = (n_p, 3, H // 8, W // 8)
    row_count = np.stack([nbytes.reshape(shp).sum(-1), esc_per_block.reshape(shp).sum(-1)], -1).astype(np.uint32)
    return np.ascontiguousarray(bitmap), row_count, nibbles, np.ascontiguousarray(vals[~small])


def check_mv_range(mv, H, W, bs):
    """Every vector must keep its macroblock inside the frame (the reference's MC copy would raise otherwise,
    motion.py:62-65); a container from the wire is not trusted."""
    nbx = W // bs
    k = np.arange(mv.shape[-2])
    x = (k % nbx) * bs + mv[..., 0].astype(np.int32)
    y = (k // nbx) * bs + mv[..., 1].astype(np.int32)
    if mv.size and (x.min() < 0 or y.min() < 0 or x.max() > W - bs or y.max() > H - bs):
        raise ValueError("container holds a motion vector that points outside the frame")


def unpack(buf) -> dict:
    """Inverse of pack(): dict(T, H, W, block_size, gop_len, coef_mode, qf, Q, i_frames, mv, coef) -- arrays are views."""
    buf = memoryview(buf)
    if len(buf) < _HDR.size:
        raise ValueError("truncated container")
    magic, ver, T, H, W, bs, gop, cm, n_p, N, qf, n_nib, n_esc = _HDR.unpack(buf[:_HDR.size])
    if magic != MAGIC or ver not in (VERSION, 2):
        raise ValueError("not a vcs_b200 container (magic / version)")
    if cm not in COEF_DTYPES or gop < 2 or bs < 1 or n_p != _capi.num_p_frames(T, gop) or N != (H // bs) * (W // bs):
        raise ValueError("inconsistent container header")
    n_i = (T + gop - 1) // gop
    dt = np.dtype(COEF_DTYPES[cm])
    if ver == 2:
        if cm != _capi.COEF_I8_RINT or H % 8 or W % 8:
            raise ValueError("inconsistent container header")
        sizes = [3 * 64 * 8, n_i * H * W * 3, n_p * N * 2 * 2, n_p * 3 * (H // 8) * (W // 8) * 8, n_p * 3 * (H // 8) * 8, n_nib, n_esc]
    else:
        sizes = [3 * 64 * 8, n_i * H * W * 3, n_p * N * 2 * 2, n_p * 3 * H * W * dt.itemsize]
    if len(buf) != _HDR.size + sum(sizes):
        raise ValueError("container size does not match its header")
    off, parts = _HDR.size, []
    for n in sizes:
        parts.append(buf[off:off + n])
        off += n
    mv = np.frombuffer(parts[2], np.int16).reshape(n_p, N, 2)
    check_mv_range(mv, H, W, bs)
    if ver == 2:
        bitmap = np.frombuffer(parts[3], np.uint64).reshape(n_p, 3, H // 8, W // 8)
        row_count = np.frombuffer(parts[4], np.uint32).reshape(n_p, 3, H // 8, 2)
        if int(row_count[..., 0].sum()) != n_nib or int(row_count[..., 1].sum()) != n_esc:
            raise ValueError("row counts do not add up to the stream lengths")
        return dict(T=T, H=H, W=W, block_size=bs, gop_len=gop, coef_mode=cm, qf=qf, version=2,
                    Q=np.frombuffer(parts[0], np.float64).reshape(3, 8, 8),
                    i_frames=np.frombuffer(parts[1], np.uint8).reshape(n_i, H, W, 3), mv=mv,
                    bitmap=bitmap, row_count=row_count, nibbles=np.frombuffer(parts[5], np.uint8),
                    escapes=np.frombuffer(parts[6], np.int8), lengths=(n_nib, n_esc))
    return dict(T=T, H=H, W=W, block_size=bs, gop_len=gop, coef_mode=cm, qf=qf,
                Q=np.frombuffer(parts[0], np.float64).reshape(3, 8, 8),
                i_frames=np.frombuffer(parts[1], np.uint8).reshape(n_i, H, W, 3),
                mv=np.frombuffer(parts[2], np.int16).reshape(n_p, N, 2),
                coef=np.frombuffer(parts[3], dt).reshape(n_p, 3, H, W))


def to_frames(c: dict):
    """(encoded_frames, ref_frames) exactly as Encoder leaves them (encoder.py:38-70): I-frames are
    Frame("I", None, None, None, t, ref_idx) with the image in ref_frames; P-frames carry mv as [dx,dy] int lists,
    r as three float64 planes (the rounded indices as floats when the clip holds indices) and c as [x,y] coords."""
    H, W, bs, g = c["H"], c["W"], c["block_size"], c["gop_len"]
    if "coef" not in c:                                   # version 2: packed indices
        c = dict(c, coef=expand_packed(c["bitmap"], c["row_count"], c["nibbles"], c["escapes"], H, W))
    coords = [[x, y] for y in range(0, H - bs + 1, bs) for x in range(0, W - bs + 1, bs)]
    frames, refs, p = [], [], 0
    for t in range(c["T"]):
        if t % g == 0:
            refs.append(np.array(c["i_frames"][t // g]))
            frames.append(Frame("I", None, None, None, t, t % g))          # encoder.py:43: frame_num % len(pattern)
        else:
            planes = [np.array(c["coef"][p, k], np.float64) for k in range(3)]
            frames.append(Frame("P", np.asarray(c["mv"][p], np.int64).tolist(), planes, coords, t, t // g))
            p += 1
    return frames, refs


def from_frames(encoded_frames, ref_frames, *, block_size, gop_len, coef_mode=_capi.COEF_F64, qf=50.0, Q=None) -> bytes:
    """pack() from the reference-style objects: Encoder.encoded_frames and Encoder.ref_frames."""
    T = len(encoded_frames)
    P = [f for f in encoded_frames if f.t == "P"]
    i_frames = np.stack([np.asarray(r, np.uint8) for r in ref_frames])
    H, W = i_frames.shape[1:3]
    N = (H // block_size) * (W // block_size)
    mv = np.asarray([f.mv for f in P], np.int16).reshape(len(P), N, 2)
    coef = np.stack([np.stack([np.asarray(pl) for pl in f.r]) for f in P]) if P else np.zeros((0, 3, H, W))
    return pack(i_frames, mv, coef.astype(COEF_DTYPES[coef_mode]), T=T, block_size=block_size, gop_len=gop_len,
                coef_mode=coef_mode, qf=qf, Q=Q)
