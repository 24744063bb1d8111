"""Seeded random sweep of the search parameter space: tiled kernel == generic kernel == oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(vcs, cur, ref, bs, lo, hi, step, slack, metric, thr, kernel):
    c = vcs._capi
    H, W = cur.shape[:2]
    p = c.me_reference_params(H, W, bs)
    p.lo, p.hi, p.step, p.slack, p.metric, p.static_thr, p.kernel = lo, hi, step, slack, metric, thr, kernel
    N = c.num_blocks(H, W, bs)
    mv = np.empty((N, 2), np.int16); cost = np.empty(N, np.uint32); fl = np.empty(N, np.uint8)
    vcs.runtime.get_context().call("vcs_me_search_host", p, cur.ctypes.data, ref.ctypes.data, mv.ctypes.data,
                                   cost.ctypes.data, fl.ctypes.data)
    return mv.astype(np.int32), cost, fl


@pytest.mark.parametrize("seed", range(24))
def test_random_search_configs(orc, seed):
    import vcs_h264_b200 as vcs
    rng = np.random.default_rng(1000 + seed)
    bs = int(rng.choice([8, 16, 16, 4]))
    tiled_ok = bs in (8, 16)
    W = int(rng.integers(2, 9)) * 16 if tiled_ok else int(rng.integers(17, 90))
    H = int(rng.integers(bs, 120))
    step = 1 if tiled_ok and rng.random() < 0.8 else int(rng.integers(1, 6))
    lo = -int(rng.integers(0, 40))
    hi = int(rng.integers(0, 40)) if rng.random() < 0.8 else lo + int(rng.integers(0, 6))   # sometimes 0 not in [lo,hi]
    slack = int(rng.integers(0, 2))
    metric = int(rng.integers(0, 2))
    thr = int(rng.choice([-1, 0, 2000, 50000]))
    kind = seed % 3
    if kind == 0:
        ref = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        cur = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    elif kind == 1:   # shifted copy + noise: meaningful minima, many near-ties
        base = rng.integers(0, 256, (H + 16, W + 16, 3), dtype=np.uint8)
        base = (base.astype(np.uint16) // 32 * 32).astype(np.uint8)
        ref = np.ascontiguousarray(base[8:8 + H, 8:8 + W])
        cur = np.ascontiguousarray(base[8 + 2:8 + 2 + H, 8 - 3:8 - 3 + W])
    else:             # low-entropy frames: exact ties everywhere
        ref = rng.integers(0, 3, (H, W, 3), dtype=np.uint8) * 100
        cur = rng.integers(0, 3, (H, W, 3), dtype=np.uint8) * 100
    omv, ocost, ofl = orc.me(cur, ref, bs, lo, hi, step, slack, metric=metric, static_thr=thr)
    for kernel in (vcs.ME_GENERIC, vcs.ME_AUTO):
        mv, cost, fl = _run(vcs, cur, ref, bs, lo, hi, step, slack, metric, thr, kernel)
        ctx = (seed, bs, H, W, lo, hi, step, slack, metric, thr, kernel)
        assert np.array_equal(fl, ofl), ctx
        assert np.array_equal(mv, omv), ctx
        assert np.array_equal(cost, ocost), ctx
