"""predict_stream on the acts64 batch for several pipeline depths / packer team sizes (ms per batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_fpga_b200 import SegmentClassifier, data
dev = torch.device("cuda:0")
graphs = [data.acts_like_graph(400, seed=b) for b in range(64)]
torch.manual_seed(0)
model = SegmentClassifier(3, 32, 4).to(dev).eval()
def run(depth, steps=150):
    for _ in model.predict_stream([graphs] * 5, depth=depth):
        pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in model.predict_stream([graphs] * steps, depth=depth):
        pass
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3
for threads in ("", "14", "12", "10", "8", "6"):
    if threads:
        os.environ["GNNSEG_PACK_THREADS"] = threads
    else:
        os.environ.pop("GNNSEG_PACK_THREADS", None)
    print("pack threads %-7s" % (threads or "default"), "  ".join("depth %d: %.3f ms" % (d, run(d)) for d in (2, 3, 4)), flush=True)
