"""Debug helper: run input/edge/node steps on cuda:0 for one golden case and print error stats."""
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
HX = torch.zeros(n, h + 4, device=dev); HX2 = torch.full((n, h + 4), 7.0, device=dev)
P = torch.zeros(n, 2 * h, device=dev); P2 = torch.full((n, 2 * h), 7.0, device=dev)
e = torch.zeros(batch.n_slots, device=dev)
st = _stream_ptr(dev)
src, dst, Xh = batch.src.cpu().long(), batch.dst.cpu().long(), batch.X.cpu()
assert L.gnnseg_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(HX), _ptr(P), st) == 0
H0 = O.sparse_input(p, Xh)
e_ref = O.sparse_edge(p, H0, src, dst)
e_in = e_ref.to(dev).contiguous()
rc = L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(HX), _ptr(e_in), h, _ptr(HX2), _ptr(P2), st)
torch.cuda.synchronize()
print("node_step rc", rc)
H1 = O.sparse_node(p, H0, e_ref, src, dst)
got = HX2.cpu()
dH = (got[:, :h] - H1).abs()
print("H' max abs err %.3e  (rows with err>1e-4: %d of %d)" % (dH.max().item(), int((dH.max(1).values > 1e-4).sum()), n))
print("X part equal:", bool(torch.equal(got[:, h:h + F], Xh)), " pad zero:", bool((got[:, h + F:] == 0).all()))
HXn = torch.cat([H1, Xh], 1)
W1 = p[O.PARAM_KEYS[2]]; D = h + F
Pref = torch.cat([HXn @ W1[:, :D].T + p[O.PARAM_KEYS[3]], HXn @ W1[:, D:].T], 1)
dP = (P2.cpu() - Pref).abs()
print("P' max abs err %.3e" % dP.max().item())
if dH.max() > 1e-4:
    bad = torch.nonzero(dH.max(1).values > 1e-4).flatten()[:10]
    print("first bad rows", bad.tolist())
    r = int(bad[0]); print("got", got[r, :8].tolist()); print("ref", H1[r, :8].tolist())
