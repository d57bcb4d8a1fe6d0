#!/usr/bin/env python
"""Where does the end-to-end stream lose time when several ranks share a host?  Under torchrun every rank streams the same
acts64 store batch through model.predict_stream and reports, per batch: wall time, host time inside the library call,
host time waiting for results (GNNSEG_STREAM_TIMING=1).  Variants: pipeline depth; only a subset of the ranks active."""
import os, sys, time
os.environ["GNNSEG_STREAM_TIMING"] = "1"
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gnn_fpga_b200 import SegmentClassifier, GraphStore
from gnn_fpga_b200 import dist as gdist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    if os.environ.get("GNNSEG_BIND_CPUS", "1") != "0":
        gdist.bind_to_local_cpus(local)
cfg = bench.WORKLOADS["acts64"]
graphs = bench.make_graphs("acts64", rank)
torch.manual_seed(0)
model = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"]).to(dev).eval()
store = GraphStore.from_sparse_graphs(graphs, reorder="none")
sb = store.batch(0, len(graphs))
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
N = 300
for active in sorted({world, max(1, world // 2), 1}, reverse=True):
    for depth in (3, 6):
        with torch.no_grad():
            for _ in model.predict_stream([sb] * 8, depth=depth):
                pass
            barrier()
            t0 = time.perf_counter()
            if rank < active:
                for _ in model.predict_stream([sb] * N, depth=depth):
                    pass
                torch.cuda.synchronize(dev)
            wall = (time.perf_counter() - t0) / N * 1e3
            st = model._stream_stats or {"call": 0, "wait": 0, "batches": 1}
            barrier()
        if rank < active:
            msg = "active %d depth %d rank %d: wall %.3f ms per batch, library call %.3f, waiting %.3f" % (
                active, depth, rank, wall, st["call"] / max(1, st["batches"]) * 1e3, st["wait"] / max(1, st["batches"]) * 1e3)
            if rank in (0, active - 1):
                print(msg, flush=True)
if world > 1:
    dist.destroy_process_group()
