"""Host <-> device copy bandwidth from pinned memory, one process per GPU (run under torchrun to load every GPU at
once): H2D alone, D2H alone, both directions together.  Rank 0 prints per-GPU and aggregate GB/s."""
import os
import time

import torch

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


rows = []
for name, fn, nbytes in (("H2D", lambda: d.copy_(h, non_blocking=True), n), ("D2H", lambda: h.copy_(d, non_blocking=True), n),
                         ("both directions", both, 2 * n)):
    fn(); torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    if dist is not None:
        x = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dt = float(x.item())
    rows.append((name, nbytes / dt / 1e9))
if rank == 0:
    for name, g in rows:
        print(f"{name:16s} {g:6.1f} GB/s per GPU, {g * world:7.1f} GB/s over {world} GPU(s)")
if dist is not None:
    dist.destroy_process_group()
