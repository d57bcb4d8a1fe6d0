"""Debug helper: run input/edge/node steps on cuda:0 for one case and print error stats."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_case
from oracle import segclf_oracle as O
from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, _lib, data
from gnn_fpga_b200.graph import _ptr, _stream_ptr

name = sys.argv[1] if len(sys.argv) > 1 else "acts_ragged_h32_it4"
dev = torch.device("cuda:0")
if name.startswith("big"):
    h = int(name[3:]) if len(name) > 3 else 32
    graphs = [data.acts_like_graph(400, seed=i) for i in range(2)]
    p = O.init_params(3, h, 0)
    model = SegmentClassifier(3, h, 2); model.load_state_dict(p); model = model.to(dev).eval()
    batch = DeviceGraphBatch.from_sparse_graphs(graphs, dev)
    F = 3
else:
    rec = load_case(name)
    model = SegmentClassifier(rec["F"], rec["h"], rec["n_iters"], masks_e=rec["masks_e"], masks_n=rec["masks_n"])
    model.load_state_dict(rec["params"]); model = model.to(dev).eval()
    batch = DeviceGraphBatch.from_dense(*[torch.from_numpy(rec[k].astype(np.float32)).to(dev) for k in ("X", "Ri", "Ro")])
    p = O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"]); h = rec["h"]; F = rec["F"]
L = _lib.lib()
n = batch.n_nodes
blob = model.pack_weights()
X4 = torch.zeros(n, 4, device=dev)
P = torch.zeros(n, 2 * h, device=dev); Q = torch.zeros(n, 3 * h, device=dev)
P2 = torch.full((n, 2 * h), 7.0, device=dev); Q2 = torch.full((n, 3 * h), 7.0, device=dev)
st = _stream_ptr(dev)
src, dst, Xh = batch.src.cpu().long(), batch.dst.cpu().long(), batch.X.cpu()
assert L.gnnseg_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(P), _ptr(Q), st) == 0
H0 = O.sparse_input(p, Xh)
Pref, Qref = O.projections(p, H0)
print("input: P err %.3e  Q err %.3e" % ((P.cpu() - Pref).abs().max().item(), (Q.cpu() - Qref).abs().max().item()))
e_ref = O.sparse_edge(p, H0, src, dst)
e_dev = e_ref.to(dev); Q_in = Qref.to(dev).contiguous()
n_in, n_out = int(batch.in_ptr[-1]), int(batch.out_ptr[-1])
e_in = e_dev[batch.in_eid[:n_in].long()].contiguous(); e_out = e_dev[batch.out_eid[:n_out].long()].contiguous()
rc = L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(X4), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h, _ptr(P2), _ptr(Q2), st)
torch.cuda.synchronize()
print("node_step rc", rc)
H1 = torch.cat([O.sparse_node(p, H0, e_ref, src, dst), Xh], 1)
P1, Q1 = O.projections(p, H1)
dP = (P2.cpu() - P1).abs(); dQ = (Q2.cpu() - Q1).abs()
print("node: P' max abs err %.3e (bad rows %d)  Q' max abs err %.3e (bad rows %d) of %d" % (
    dP.max().item(), int((dP.max(1).values > 1e-4).sum()), dQ.max().item(), int((dQ.max(1).values > 1e-4).sum()), n))
if dP.max() > 1e-4:
    bad = torch.nonzero(dP.max(1).values > 1e-4).flatten()[:10]
    print("first bad rows", bad.tolist()); r = int(bad[0]); print("got", P2[r, :8].tolist()); print("ref", P1[r, :8].tolist())
# end-to-end accuracy of the full forward vs fp32 / fp64 oracles
if name.startswith("big"):
    out = model(graphs).cpu().numpy()
    Xf, s_, d_, _ = O.flatten_sparse_batch(graphs)
    for dt in (torch.float32, torch.float64):
        ref = O.sparse_forward(p, Xf, s_, d_, 2, dt).numpy().reshape(out.shape)
        print("forward rel err vs", dt, float(np.max(np.abs(out - ref) / np.abs(ref))))
