"""Does L2 residency matter?  Times edge / node steps back to back on the same inputs, for several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, _lib
from gnn_fpga_b200.graph import _ptr, _stream_ptr

dev = torch.device("cuda:0")
L = _lib.lib()
cfg = bench.WORKLOADS["acts64"]
h, F = cfg["h"], cfg["F"]
graphs_all = bench.make_graphs("acts64", 0)
torch.manual_seed(0)
model = SegmentClassifier(F, h, 4).to(dev).eval()
blob = model.pack_weights()
st = _stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for nb in (64, 32, 16, 8):
    batch = DeviceGraphBatch.from_sparse_graphs(graphs_all[:nb], dev)
    n, m = batch.n_nodes, batch.n_slots
    X4 = torch.empty(n, 4, device=dev); P = torch.empty(n, 2 * h, device=dev)
    Q = [torch.empty(n, 3 * h, device=dev) for _ in range(2)]
    e = torch.empty(m, device=dev); e_in = torch.empty(m, device=dev); e_out = torch.empty(m, device=dev)
    def ev(): return torch.cuda.Event(enable_timing=True)
    res = {}
    for rep in range(3):
        flush.zero_()
        seq = []
        def run(name, fn):
            a, b = ev(), ev(); a.record(); rc = fn(); b.record(); assert rc == 0; seq.append((name, a, b))
        run("input", lambda: L.gnnseg_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(P), _ptr(Q[0]), st))
        for k in range(3):
            run("edge%d" % k, lambda: L.gnnseg_edge_step(_ptr(blob), C.byref(batch.struct), _ptr(P), h, None, _ptr(e_in), _ptr(e_out), st))
        for k in range(3):
            run("node%d" % k, lambda: L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(X4), _ptr(Q[0]), _ptr(e_in), _ptr(e_out), h, _ptr(P), _ptr(Q[1]), st))
        run("edgeA", lambda: L.gnnseg_edge_step(_ptr(blob), C.byref(batch.struct), _ptr(P), h, None, _ptr(e_in), _ptr(e_out), st))
        run("edge_slot", lambda: L.gnnseg_edge_step(_ptr(blob), C.byref(batch.struct), _ptr(P), h, _ptr(e), None, None, st))
        torch.cuda.synchronize()
        res = {nm: a.elapsed_time(b) * 1e3 for nm, a, b in seq}
    print("events %2d nodes %6d slots %7d : %s" % (nb, n, m, "  ".join("%s %.1f" % (k, v) for k, v in res.items())), flush=True)
