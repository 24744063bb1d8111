"""GOP sharding across ranks.

P-frames depend only on the ORIGINAL I-frame of their GOP (encoder.py:42,51-52), so GOPs are
independent units: rank r takes a contiguous range of GOPs, there is no data-path collective,
and results are gathered once at the end (torch.distributed, NCCL on GPUs / gloo in CPU tests).
"""
from __future__ import annotations


def gop_range(num_gops: int, rank: int, world: int):
    """Contiguous, balanced split: the first num_gops % world ranks get one GOP more."""
    base, extra = divmod(num_gops, world)
    g0 = rank * base + min(rank, extra)
    return g0, g0 + base + (1 if rank < extra else 0)


def frame_range(T: int, gop_len: int, rank: int, world: int):
    """Frames [t0, t1) of rank's GOPs; always starts on an I-frame."""
    num_gops = (T + gop_len - 1) // gop_len
    g0, g1 = gop_range(num_gops, rank, world)
    return min(g0 * gop_len, T), min(g1 * gop_len, T)


def p_count(T: int, gop_len: int) -> int:
    return T - (T + gop_len - 1) // gop_len if T > 0 else 0


def gather_p_outputs(local, T, gop_len, dist, device=None):
    """all_gather per-P-frame tensors (dim 0 = local P-frame ordinal) into clip order.

    local: dict name -> tensor [nP_local, ...].  Shards may hold different counts, so tensors are
    padded to the maximum before the collective and trimmed after.  Returns dict name -> tensor
    [nP_total, ...] on every rank."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = []
    for r in range(world):
        t0, t1 = frame_range(T, gop_len, r, world)
        counts.append(p_count(t1 - t0, gop_len))
    mx = max(counts) if counts else 0
    out = {}
    for name, t in local.items():
        if t is None:
            continue
        assert t.shape[0] == counts[rank], (name, t.shape, counts[rank])
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        raw = pad.reshape(-1).view(torch.uint8)       # byte payload: any dtype, NCCL or gloo
        bufs = [torch.empty_like(raw) for _ in range(world)]
        dist.all_gather(bufs, raw)
        out[name] = torch.cat([b.view(t.dtype).reshape(pad.shape)[:c] for b, c in zip(bufs, counts)], 0)
    return out


class GatherPlan:
    """all_gather of per-P-frame tensors with buffers allocated ONCE (bench.py's C3 leg gathers gigabytes per step; the
    pad/cat temporaries of gather_p_outputs would dominate).  gather() returns per-rank views in clip order."""

    def __init__(self, like, T, gop_len, dist):
        import torch
        self.dist, self.world, self.rank = dist, dist.get_world_size(), dist.get_rank()
        self.counts = []
        for r in range(self.world):
            t0, t1 = frame_range(T, gop_len, r, self.world)
            self.counts.append(p_count(t1 - t0, gop_len))
        mx = max(self.counts)
        self.shapes, self.recv, self.send = {}, {}, {}
        for name, t in like.items():
            per = int(t[0].numel()) * t.element_size() if t.shape[0] else 0
            self.shapes[name] = (tuple(t.shape[1:]), t.dtype, per)
            self.recv[name] = torch.empty((self.world, mx * per), dtype=torch.uint8, device=t.device)
            self.send[name] = torch.zeros(mx * per, dtype=torch.uint8, device=t.device) if self.counts[self.rank] < mx else None

    def bytes_per_step(self):
        return sum(int(self.recv[n].numel()) for n in self.recv)

    def gather(self, local):
        out = {}
        for name, t in local.items():
            shape, dtype, per = self.shapes[name]
            raw = t.reshape(-1).view(self.dist_uint8())
            if self.send[name] is not None:          # shorter shard: pad to the common length
                self.send[name][:raw.numel()] = raw
                raw = self.send[name]
            self.dist.all_gather_into_tensor(self.recv[name].view(-1), raw)
            out[name] = [self.recv[name][r, :c * per].view(dtype).reshape((c,) + shape) for r, c in enumerate(self.counts)]
        return out

    @staticmethod
    def dist_uint8():
        import torch
        return torch.uint8
