import sys, time
sys.path.insert(0, ".")
import torch
from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, pack_sparse_batch_host, data
dev = torch.device("cuda:0")
graphs = [data.acts_like_graph(400, seed=b) for b in range(64)]
torch.manual_seed(0)
model = SegmentClassifier(3, 32, 4).to(dev).eval()
def T(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r
ms, host = T(lambda: pack_sparse_batch_host(graphs)); print("host pack (alloc pinned each time) %.3f ms" % ms)
pin = {"X": torch.empty((300000, 3), pin_memory=True), "src": torch.empty(1500000, dtype=torch.int32, pin_memory=True), "dst": torch.empty(1500000, dtype=torch.int32, pin_memory=True)}
ms, host = T(lambda: pack_sparse_batch_host(graphs, pinned=pin)); print("host pack (reused pinned)        %.3f ms" % ms)
for nt in (1, 4, 16, 64):
    ms, _ = T(lambda: pack_sparse_batch_host(graphs, pinned=pin, n_threads=nt)); print("  threads=%d %.3f ms" % (nt, ms))
ms, d = T(lambda: [host[k].to(dev, non_blocking=True) for k in ("X", "src", "dst")]); print("H2D %.3f ms" % ms)
X, src, dst = d
ms, batch = T(lambda: DeviceGraphBatch(X, src, dst, 64, host["e_max"])); print("DeviceGraphBatch (CSR build + allocs) %.3f ms" % ms)
model.use_cuda_graph = False
ms, out = T(lambda: model(batch)); print("forward (no graph) %.3f ms" % ms)
ms, _ = T(lambda: out.cpu()); print("D2H pageable %.3f ms" % ms)
hp = torch.empty(out.shape, pin_memory=True)
ms, _ = T(lambda: hp.copy_(out, non_blocking=True)); print("D2H pinned %.3f ms" % ms)
ms, _ = T(lambda: model(graphs).cpu()); print("e2e total %.3f ms" % ms)
