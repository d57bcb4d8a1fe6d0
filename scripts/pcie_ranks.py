#!/usr/bin/env python
"""(Round 2: written for a last diagnostic, whose first version had one barrier more on the active ranks than on the
idle ones and hung -- it took the round's remaining GPU minutes with it; this version is fixed but has not been run.)
Host <-> device copy rates when every rank of a node copies at once (torchrun): the e2e stream moves 10.3 MB to the
device and 5.4 MB back per acts64 batch.  Per rank: ms per iteration for H2D alone, D2H alone, both on two streams."""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H2D, D2H, N = 10_287_168, 5_435_908, 300
src = torch.empty(H2D, dtype=torch.uint8, pin_memory=True); src.fill_(1)
dst = torch.empty(D2H, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev)
d_out = torch.zeros(D2H, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
def run(h2d, d2h):
    t0 = time.perf_counter()
    for _ in range(N):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(src, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                dst.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / N * 1e3
res = []
for active in sorted({world, max(1, world // 2), 1}, reverse=True):
    for h2d, d2h in ((True, False), (False, True), (True, True)):
        barrier()
        t = run(h2d, d2h) if rank < active else 0.0
        barrier()
        res.append((active, h2d, d2h, t))
out = torch.tensor([r[3] for r in res], device=dev)
if world > 1:
    allr = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(allr, out)
else:
    allr = [out]
if rank == 0:
    for i, (active, h2d, d2h, _) in enumerate(res):
        ts = [float(a[i]) for a in allr[:active]]
        gb = (H2D * h2d + D2H * d2h) / 1e6
        print("active %d %s: ms per iteration by rank %s  -> slowest %.1f GB/s per GPU, %.0f GB/s in total" % (
            active, ("H2D" if h2d else "") + ("+" if h2d and d2h else "") + ("D2H" if d2h else ""),
            " ".join("%.3f" % t for t in ts), gb / max(ts), gb * active / max(ts)), flush=True)
if world > 1:
    dist.destroy_process_group()
