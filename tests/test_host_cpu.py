"""CPU-only checks: the C-ABI library loads and exports every declared symbol, its pure host
functions agree with the reference's golden values, the host-side mirror keeps the reference's
surface, and GOP sharding / gather logic (world_size 2, gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from vcs_h264_b200 import _capi
    _capi.load()
    return _capi


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "vcs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vcs_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)


def test_ctypes_signatures_follow_the_header(capi):
    """Every prototype of include/vcs_b200.h against its ctypes binding: same number of parameters, and pointers /
    integers / doubles / 64-bit sizes in the same places (a drifted binding would pass garbage through the C ABI)."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "vcs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\b([a-z_A-Z0-9 ]+?[ \*]+)(vcs_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)
    assert len(protos) >= 30

    def kind_of_c(decl):
        decl = decl.strip()
        if "*" in decl or "[" in decl:
            return "ptr"
        base = re.sub(r"\b(const|unsigned|signed)\b", "", decl).split()
        t = base[0] if base else "int"
        if t == "double":
            return "f64"
        if t in ("size_t", "int64_t", "uint64_t") or decl.startswith("unsigned long long") or decl.startswith("long long"):
            return "i64"
        return "i32"

    def kind_of_ctypes(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        if t is C.c_double:
            return "f64"
        return "i64" if C.sizeof(t) == 8 else "i32"

    seen = set()
    for ret, name, params in protos:
        seen.add(name)
        plist = [q for q in (x.strip() for x in params.split(",")) if q and q != "void"]
        res, args = capi.SIGNATURES[name]
        assert len(plist) == len(args), (name, plist, args)
        for decl, t in zip(plist, args):
            assert kind_of_c(decl) == kind_of_ctypes(t), (name, decl, t)
        assert kind_of_c(ret + "x") == kind_of_ctypes(res), (name, ret, res)
    assert seen == set(capi.SIGNATURES)


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.VcsError):
        capi.Context(0)


def test_product_never_imports_oracle():
    """The product package may mention the oracle in comments but never import, link or call it."""
    pkg = os.path.join(ROOT, "vcs_h264_b200")
    bad = re.compile(r"^\s*(from|import)\s+\S*oracle|libvcs_oracle|vcs_oracle_\w+\s*\(|oracle\.(me|compress|"
                     r"encode_p|decompress)\(|#include\s+\S*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f


def test_host_constants_match_reference(capi, golden):
    assert np.array_equal(capi.dct_matrix(), golden["dctmat"])          # DCTcompressor.py:124-133
    for qf in (1, 10, 49, 50, 75, 99):
        assert np.array_equal(capi.q_tables(float(qf)), golden[f"Q_{qf}"])  # DCTcompressor.py:29-38
    with pytest.raises(ValueError):
        capi.q_tables(100.0)


def test_reference_params(capi, orc):
    for bs in (4, 8, 16):
        p = capi.me_reference_params(360, 640, bs)
        o = orc.reference_search_params(bs)
        assert (p.lo, p.hi, p.step, p.slack) == (o["lo"], o["hi"], o["step"], o["slack"])
        assert p.metric == capi.METRIC_WRAP8 and p.static_thr == 2000
    p = capi.me_fullsearch_params(1080, 1920, 16, 16)
    assert (p.lo, p.hi, p.step, p.slack) == (-16, 16, 1, 0)
    assert capi.num_blocks(1080, 1920, 16) == 8040                         # C2
    assert capi.num_blocks(360, 640, 8) == 3600
    assert capi.num_p_frames(60, 4) == 45 and capi.num_p_frames(152, 4) == 114
    assert capi.num_p_frames(5, 4) == 3 and capi.num_p_frames(1, 4) == 0


def test_dropin_surface():
    import vcs_h264_b200 as v
    from vcs_h264_b200 import motion
    assert (motion.SIMILARITY_THRESHOLD, motion.CHANNELS, motion.WRITE_STATIC_BLOCK) == (2000, 3, True)
    mp = v.MotionProcessor(8, [360, 640])
    assert (mp.block_size, mp.shape, mp.search_window_size) == (8, [360, 640], 16)
    blocks, coords = mp._split_frame_into_mblocks(np.zeros((360, 640, 3), np.uint8))
    assert len(coords) == 3600 and coords[0] == [0, 0] and coords[-1] == [632, 352]
    assert blocks[1].shape == (8, 8, 3)
    assert mp._get_motion_vector([5, 9], [8, 8]) == [-3, 1]
    mp2 = v.MotionProcessor(16, [1080, 1920])
    assert len(mp2._block_coords()) == 8040
    dc = v.DCTCompressor(8)
    assert dc.blocksize == 8 and dc.compressed == [] and len(dc.Q) == 3
    f = v.Frame("P", [[0, 0]], None, [[0, 0]], 3, 0)
    assert (f.t, f.mv, f.r, f.c, f.i, f.ref_i) == ("P", [[0, 0]], None, [[0, 0]], 3, 0)
    enc = v.Encoder(["I", "P", "P", "P"], [360, 640], 8, True)
    assert enc.ENCODING_PATTERN_LENGTH == 4 and enc.ref_frames == [] and enc.encoded_frames == []
    with pytest.raises(ValueError):
        v.DCTCompressor(16).compress(np.zeros((16, 16, 3), np.uint8))
    # private helpers and attributes the reference's classes carry (DCTcompressor.py:100-139, decoder.py:16)
    for name in ("_dct2", "_idct2", "_dctMatrix", "_cuHelper", "_completeDCT", "compress", "decompress"):
        assert callable(getattr(dc, name)), name
    assert dc._cuHelper(0) == 2 ** -0.5 and dc._cuHelper(3) == 1
    from vcs_h264_b200 import DCTcompressor as dmod
    assert (dmod.QF, dmod.TEST_COMPRESSOR, dmod.QUANTIZE) == (50.0, False, False) and len(dmod.Q) == 3
    dec = v.Decoder([], 25, [360, 640], [], 8, True)
    assert dec.fourcc == 875967064                       # cv2.VideoWriter_fourcc(*'X264')
    for name in ("reconstruct_video", "_fully_reconstruct", "_reconstruct_P_frame"):
        assert callable(getattr(dec, name)), name
    for name in ("encode_frame", "_process_I_frame", "_process_B_frame", "_process_P_frame"):
        assert callable(getattr(enc, name)), name


def test_import_has_no_side_effects():
    """Importing the package (or its synth module) neither dlopens the CUDA library nor compiles anything;
    bench.py's reference arm relies on that."""
    code = ("import sys; sys.path.insert(0, %r); import vcs_h264_b200, vcs_h264_b200.synth, vcs_h264_b200.DCTcompressor;"
            "maps = open('/proc/self/maps').read(); assert 'libvcs_b200' not in maps, 'library mapped at import';"
            "from vcs_h264_b200 import _capi; assert _capi._lib is None; print('clean')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout + r.stderr


def test_sharding_ranges():
    from vcs_h264_b200 import sharding
    for T, g in ((60, 4), (240, 4), (18, 4), (7, 3), (1, 4)):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.frame_range(T, g, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            assert all(t0 % g == 0 for t0, t1 in spans if t1 > t0)
            assert sum(sharding.p_count(t1 - t0, g) for t0, t1 in spans) == sharding.p_count(T, g)
    assert sharding.gop_range(60, 3, 8) == (24, 32)


def test_synth_clip_is_deterministic_and_non_static(orc):
    from vcs_h264_b200 import synth
    a = synth.clip(3, 64, 96, seed=1, margin=32)
    b = synth.clip(3, 64, 96, seed=1, margin=32)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    mv, cost, flags = orc.me(a[1], a[0], 16, **orc.symmetric_search_params(16))
    assert not np.any(flags & 1)


def test_gather_world2_gloo(tmp_path):
    """gather_p_outputs over 2 gloo ranks reassembles per-shard outputs in clip order."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from vcs_h264_b200 import sharding
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
T, g = 22, 4
t0, t1 = sharding.frame_range(T, g, r, w)
ps = [t for t in range(t0, t1) if t % g]
local = dict(mv=torch.tensor(ps, dtype=torch.int16).reshape(-1, 1, 1).repeat(1, 5, 2),
             cost=torch.tensor(ps, dtype=torch.int32).reshape(-1, 1).repeat(1, 5))
out = sharding.gather_p_outputs(local, T, g, dist)
want = [t for t in range(T) if t % g]
assert out["mv"][:, 0, 0].tolist() == want, out["mv"][:, 0, 0]
assert out["cost"][:, 3].tolist() == want
dist.destroy_process_group()
print("OK", r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29511")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


def test_gather_plan_world2_gloo(tmp_path):
    """GatherPlan (bench.py's C3 leg): preallocated all_gather of unequal shards, per-rank views in clip order."""
    script = tmp_path / "w2.py"
    script.write_text(f"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
from vcs_h264_b200 import sharding
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
T, g = 22, 4                                   # 6 GOPs (the last one short): 3 + 3 GOPs, 9 + 7 P-frames
t0, t1 = sharding.frame_range(T, g, r, w)
ps = [t for t in range(t0, t1) if t % g]
local = dict(mv=torch.tensor(ps, dtype=torch.int16).reshape(-1, 1, 1).repeat(1, 5, 2).contiguous(),
             coef=torch.tensor(ps, dtype=torch.int8).reshape(-1, 1, 1, 1).repeat(1, 3, 4, 8).contiguous())
plan = sharding.GatherPlan(local, T, g, dist)
for _ in range(2):                             # buffers are reused
    out = plan.gather(local)
want = [t for t in range(T) if t % g]
got = torch.cat(out["mv"], 0)[:, 0, 0].tolist()
assert got == want, got
assert torch.cat(out["coef"], 0)[:, 2, 3, 7].tolist() == want
assert [x.shape[0] for x in out["mv"]] == plan.counts and sum(plan.counts) == len(want)
dist.destroy_process_group()
print("OK", r)
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29513")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29513", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


def test_only_tests_bench_and_smoke_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package or tools/ may import it."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|.*libvcs_oracle)", re.M)
    for f in glob.glob(os.path.join(root, "tools", "*.py")) + glob.glob(os.path.join(root, "vcs_h264_b200", "**", "*.py"), recursive=True):
        assert not pat.search(open(f).read()), f
