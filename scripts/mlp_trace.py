"""clock64 stamps per role of CTA 0 of node_mlp_kernel_tc<32> and %globaltimer per CTA.
    make -C gnn_fpga_b200/csrc trace
    GNNSEG_LIB=gnn_fpga_b200/libgnnseg_trace.so python scripts/mlp_trace.py
(the instrumented library is a separate build, see the Makefile; the shipped one carries no stamps)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_fpga_b200 import _lib, data, DeviceGraphBatch, SegmentClassifier
dev = torch.device("cuda:0")
graphs = [data.acts_like_graph(400, seed=b) for b in range(64)]
batch = DeviceGraphBatch.from_sparse_graphs(graphs, device=dev)
torch.manual_seed(0)
model = SegmentClassifier(3, 32, 4).to(dev).eval()
model.use_cuda_graph = False
L = _lib.lib()
blob = model.pack_weights()
n = batch.n_nodes
X4 = torch.randn(n, 4, device=dev); h1 = torch.tanh(torch.randn(n, 32, device=dev))
P = torch.empty(n, 64, device=dev); Q = torch.empty(n, 96, device=dev)
ptr = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):      # a node MLP step that writes P' and Q' (the common case: every node step but the last)
    _lib.check(L.gnnseg_node_mlp_step(ptr(blob), ptr(X4), ptr(h1), 32, n, 32, ptr(P), ptr(Q), None), "node_mlp_step")
torch.cuda.synchronize()
buf = (C.c_longlong * (2 * 16 * 12))()
assert L.gnnseg_debug_read_trace(buf) == 0
t = np.array(buf, dtype=np.int64).reshape(2, 16, 12)
mlp, ldr = t[0], t[1]
names = ["tile start", "after FULL", "GEMM2 done", "epi2 done", "after EPI", "GEMM3 done", "stores issued", "after EPI 2"]
print("MLP warp 0 thread 0, CTA 0, gnnseg_node_mlp_step writing P' and Q'; cycles since the tile's start")
for it in range(8):
    if mlp[it, 0] == 0:
        break
    print("tile %d:" % it, "  ".join("%s +%d" % (names[k], mlp[it, k] - mlp[it, 0]) for k in range(1, 8)),
          " | next tile starts +%d" % (mlp[it + 1, 0] - mlp[it, 0] if mlp[it + 1, 0] else -1))
print("loader warp 0 thread 0: wait EMPTY, write A + arrive FULL, issue next fetch")
for it in range(8):
    if ldr[it, 0] == 0:
        break
    print("tile %d: EMPTY wait %d, split + stores %d, fetch issue %d, (loader start - MLP tile start %d)" %
          (it, ldr[it, 1] - ldr[it, 0], ldr[it, 2] - ldr[it, 1], ldr[it, 3] - ldr[it, 2], ldr[it, 0] - mlp[it, 0]))
e = mlp[15]
print("CTA 0: entry -> copies issued %d, weights arrived %d, prologue done %d, first tile starts %d; last tile end -> CTA done %d; entry -> CTA done %d cycles"
      % (e[1] - e[0], e[2] - e[0], e[3] - e[0], mlp[0, 0] - e[0], e[5] - e[4], e[5] - e[0]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    _lib.check(L.gnnseg_node_mlp_step(ptr(blob), ptr(X4), ptr(h1), 32, n, 32, ptr(P), ptr(Q), None), "node_mlp_step")
b.record(); torch.cuda.synchronize()
print("back-to-back launches (warm L2): %.2f us each" % (a.elapsed_time(b) * 1e3 / 20))
cb = (C.c_ulonglong * 1024)()
assert L.gnnseg_debug_read_cta(cb) == 0
ct = np.array(cb, dtype=np.uint64).reshape(512, 2)[:296].astype(np.int64)
t0 = ct[:, 0].min()
st, en = (ct[:, 0] - t0) / 1e3, (ct[:, 1] - t0) / 1e3
print("last launch, %%globaltimer: CTA starts %.2f .. %.2f us (median %.2f), ends %.2f .. %.2f us (median %.2f), lifetime %.2f .. %.2f us"
      % (st.min(), st.max(), np.median(st), en.min(), en.max(), np.median(en), (en - st).min(), (en - st).max()))
order = np.argsort(en)
print("latest CTAs:", [(int(i), round(float(st[i]), 2), round(float(en[i]), 2)) for i in order[-5:]])
