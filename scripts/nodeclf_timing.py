"""Time NodeClassifier (per-hit model) beside SegmentClassifier on the bench workloads:
inference forward and one autograd step (forward + BCELoss + backward), batch resident, CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_fpga_b200 import data, DeviceGraphBatch, SegmentClassifier
from gnn_fpga_b200.node_classifier import NodeClassifier

dev = torch.device("cuda:0")


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, graphs, h, T in (("acts64", [data.acts_like_graph(400, seed=b) for b in range(64)], 32, 4),
                           ("mu200", [data.mu200_like_graph(seed=0)], 64, 8)):
    batch = DeviceGraphBatch.from_sparse_graphs(graphs, device=dev)
    torch.manual_seed(0)
    for cls in (SegmentClassifier, NodeClassifier):
        model = cls(3, h, T).to(dev)
        model.use_cuda_graph = False
        model.eval()
        with torch.no_grad():
            t_fwd = timed(lambda: model(batch))
        model.train()
        out = model(batch)
        y = (torch.rand_like(out) < 0.3).float()

        def step():
            model.zero_grad()
            o = model(batch)
            torch.nn.functional.binary_cross_entropy(o, y).backward()
        t_step = timed(step, reps=10, warm=3)
        print("%-7s %-17s nodes %7d slots %8d: forward %.3f ms (plain launches), forward + BCE + backward %.3f ms"
              % (name, cls.__name__, batch.n_nodes, batch.n_slots, t_fwd, t_step))
