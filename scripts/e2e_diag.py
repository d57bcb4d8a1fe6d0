#!/usr/bin/env python
"""Wall-clock timeline of model.predict_stream over store batches: per-yield intervals for several depths
(is the pipeline overlapping H2D / compute / D2H, what does the first call of a stream cost)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gnn_fpga_b200 import SegmentClassifier, GraphStore

wl = sys.argv[1] if len(sys.argv) > 1 else "acts64"
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
graphs = bench.make_graphs(wl, 0)
torch.manual_seed(0)
model = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"]).to(dev).eval()
store = GraphStore.from_sparse_graphs(graphs, reorder="none")
sb = store.batch(0, len(graphs))
for depth in (3, 1, 2, 3, 4):
    for n in (3, 30, 30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        stamps = []
        for out in model.predict_stream([sb] * n, depth=depth):
            stamps.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        d = np.diff([0.0] + stamps) * 1e3
        print("depth %d n %2d total %.2f ms  per batch %.3f ms  first yield %.2f ms  median gap %.3f  max gap %.3f" %
              (depth, n, total * 1e3, total / n * 1e3, stamps[0] * 1e3, float(np.median(d[1:])) if n > 1 else 0.0, float(d[1:].max()) if n > 1 else 0.0), flush=True)
