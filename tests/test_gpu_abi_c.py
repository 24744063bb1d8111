"""The C ABI from a plain C program (gcc + dlopen, no Python in the callee): same numbers as the oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frames():
    H, W = 64, 96
    s = np.uint32(2463534242)
    ref = np.empty(H * W * 3, np.uint8)
    v = int(s)
    for i in range(ref.size):
        v ^= (v << 13) & 0xFFFFFFFF; v ^= v >> 17; v ^= (v << 5) & 0xFFFFFFFF
        ref[i] = (v >> 11) & 0xFF
    ref = ref.reshape(H, W, 3)
    ys = np.minimum(np.arange(H) + 2, H - 1); ys[np.arange(H) + 2 >= H] = np.arange(H)[np.arange(H) + 2 >= H]
    xs = np.where(np.arange(W) >= 3, np.arange(W) - 3, np.arange(W))
    cur = np.ascontiguousarray(ref[ys][:, xs])
    return cur, ref


def test_c_client(tmp_path, orc):
    from vcs_h264_b200 import _capi
    _capi.load()
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(ROOT, "tests", "abi_smoke.c"), "-ldl"])
    out = subprocess.run([exe, _capi.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    cur, ref = _frames()

    def sums(mv, cost, flags):
        k = np.arange(len(cost), dtype=np.int64)
        smv = int((mv[:, 0].astype(np.int64) * 131 + mv[:, 1] * 7 + k * (mv[:, 0].astype(np.int64) ^ mv[:, 1])).sum())
        return smv, int((cost.astype(np.int64) % 1000003).sum()), int(flags.sum())
    mv0, c0, f0 = orc.me(cur, ref, 8, **orc.reference_search_params(8))
    mv1, c1, f1 = orc.me(cur, ref, 16, metric=orc.METRIC_SAD, static_thr=-1, **orc.symmetric_search_params(8))
    assert lines[0] == "pass 0 N %d mvsum %d costsum %d flagsum %d" % ((len(c0),) + sums(mv0, c0, f0))
    assert lines[1] == "pass 1 N %d mvsum %d costsum %d flagsum %d" % ((len(c1),) + sums(mv1, c1, f1))
    pred = orc.mc(ref, 16, mv1)
    planes = orc.compress(orc.residual(cur, pred))
    outimg = orc.add_wrap(pred, orc.decompress(planes))
    ps = float((planes.ravel() * ((np.arange(planes.size) % 17) + 1)).cumsum()[-1]) if False else None
    acc = 0.0
    w = (np.arange(planes.size) % 17) + 1
    for a, b in zip(planes.ravel().tolist(), w.tolist()):   # same left-to-right double accumulation as the C loop
        acc += a * b
    osum = int((outimg.ravel().astype(np.int64) * ((np.arange(outimg.size) % 13) + 1)).sum())
    assert lines[2] == "planesum %.17g outsum %d" % (acc, osum)
