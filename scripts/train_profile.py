"""One warm NativeTrainer step on the acts64 batch (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch
from gnn_fpga_b200.training import NativeTrainer

wl = sys.argv[1] if len(sys.argv) > 1 else "acts64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
graphs = bench.make_graphs(wl, 0)
torch.manual_seed(0)
model = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"]).to(dev).train()
batch = DeviceGraphBatch.from_sparse_graphs(graphs, dev)
y = torch.zeros((len(graphs), batch.e_max))
for b, g in enumerate(graphs):
    y[b, :g.y.shape[0]] = torch.from_numpy(g.y)
y = y.to(dev)
tr = NativeTrainer(model, l1=1e-4)
for _ in range(steps):
    tr.step(batch, y)
torch.cuda.synchronize()
print("loss", float(tr.loss.item()))
