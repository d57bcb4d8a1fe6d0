"""Timing probe for the fused gather (results are NOT checked): how long does gnnseg_fused_gather_step take on the
acts64 / mu200 batch with the library given by GNNSEG_LIB?  Used to compare a build whose second row load of a visit
reads the line next to the first one (-DGNNSEG_ADJ_HACK) with the shipped one.
    GNNSEG_LIB=gnn_fpga_b200/libgnnseg_adjhack.so python scripts/gather_adjacency_probe.py [acts64|mu200]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_fpga_b200 import _lib, data, DeviceGraphBatch, GraphStore, SegmentClassifier
wl = sys.argv[1] if len(sys.argv) > 1 else "acts64"
dev = torch.device("cuda:0")
if wl == "acts64":
    graphs, h = [data.acts_like_graph(400, seed=b, edges_per_hit=5.0) for b in range(64)], 32
else:
    graphs, h = [data.acts_like_graph(10000, seed=0, edges_per_hit=10.0)], 64
batch = DeviceGraphBatch.from_store(GraphStore.from_sparse_graphs(graphs, reorder="none"), 0, len(graphs), dev)
torch.manual_seed(0)
model = SegmentClassifier(3, h, 4).to(dev).eval()
with torch.no_grad():
    model(batch)                      # builds the adjacency lists
blob = model.pack_weights()
n = batch.n_nodes
S = torch.rand(n + 1, 5 * h, device=dev) * 0.5 + 0.75
h1 = torch.empty(n + 1, 5 * h, device=dev)
L = _lib.lib()
ptr = lambda t: C.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(25):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(L.gnnseg_fused_gather_step(ptr(blob), C.byref(batch.struct), ptr(S), h, ptr(h1), 5 * h, None), "fused_gather_step")
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts = sorted(ts[5:])
print("%s %s: fused gather %.1f us (median of 20, L2 flushed), min %.1f" % (os.environ.get("GNNSEG_LIB", "default"), wl, ts[len(ts) // 2], ts[0]))
