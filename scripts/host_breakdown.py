import sys, time
sys.path.insert(0, ".")
import torch
from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, pack_sparse_batch_host, data
dev = torch.device("cuda:0")
graphs = [data.acts_like_graph(400, seed=b) for b in range(64)]
torch.manual_seed(0)
model = SegmentClassifier(3, 32, 4).to(dev).eval()
model.use_cuda_graph = False
pin = model._grow_pinned(None, graphs)
out_pin = torch.empty(64 * 25000, pin_memory=True)
N = 30
tm = {"pack": 0, "h2d": 0, "batch": 0, "run": 0, "d2h": 0}
with torch.no_grad():
    for i in range(N + 3):
        if i == 3:
            torch.cuda.synchronize(); tm = {k: 0 for k in tm}; t_all = time.perf_counter()
        t0 = time.perf_counter(); host = pack_sparse_batch_host(graphs, pinned=pin); t1 = time.perf_counter()
        X = host["X"].to(dev, non_blocking=True); s = host["src"].to(dev, non_blocking=True); d = host["dst"].to(dev, non_blocking=True); t2 = time.perf_counter()
        batch = DeviceGraphBatch(X, s, d, 64, host["e_max"]); t3 = time.perf_counter()
        sc = model._run(batch); t4 = time.perf_counter()
        out_pin[:sc.numel()].copy_(sc, non_blocking=True); t5 = time.perf_counter()
        tm["pack"] += t1 - t0; tm["h2d"] += t2 - t1; tm["batch"] += t3 - t2; tm["run"] += t4 - t3; tm["d2h"] += t5 - t4
    host_total = time.perf_counter() - t_all
    torch.cuda.synchronize()
    total = time.perf_counter() - t_all
print("host ms per batch:", {k: round(v / N * 1e3, 3) for k, v in tm.items()}, "host total", round(host_total / N * 1e3, 3), "wall", round(total / N * 1e3, 3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with torch.no_grad():
    for i in range(10):
        batch = DeviceGraphBatch(X, s, d, 64, host["e_max"]); sc = model._run(batch)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
