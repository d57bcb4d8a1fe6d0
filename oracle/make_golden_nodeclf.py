"""Generate tests/golden/nodeclf_*.npz by RUNNING THE REFERENCE'S NodeClassifier (TEST INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_nodeclf.py

The model lives in a notebook, gnn/MPNN_HitClassifier.ipynb: cell 20 defines its EdgeNetwork /
NodeNetwork, cell 21 the NodeClassifier.  This script reads those two cells' sources out of the
notebook file in the build container and executes them unmodified (nothing is copied into the
repository), builds seeded batches, and records inputs, state_dict, the model's outputs and the
loss / gradients of one nn.BCELoss() backward (what the notebook's Estimator runs, cell 29).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

NB = "/root/reference/gnn/MPNN_HitClassifier.ipynb"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from gnn_fpga_b200 import data                       # noqa: E402  (input generators only)
from oracle.make_golden import merge_graphs, ref_graph   # noqa: E402

torch.set_num_threads(1)


def reference_classes():
    cells = json.load(open(NB))["cells"]
    ns = {"torch": torch, "nn": nn, "np": np}
    found = 0
    for c in cells:
        src = "".join(c["source"])
        if c["cell_type"] == "code" and ("class EdgeNetwork" in src or "class NodeClassifier" in src):
            exec(compile(src, NB, "exec"), ns)
            found += 1
    assert found == 2, found
    return ns["NodeClassifier"]


def run_case(NodeClassifier, name, sparse_graphs, F, h, n_iters, seed):
    dense = [ref_graph.graph_from_sparse(ref_graph.SparseGraph(*g)) for g in sparse_graphs]
    X, Ri, Ro = merge_graphs(dense)
    torch.manual_seed(seed)
    net = NodeClassifier(input_dim=F, hidden_dim=h, n_iters=n_iters)
    inputs = [torch.from_numpy(X.astype(np.float32)), torch.from_numpy(Ri.astype(np.float32)),
              torch.from_numpy(Ro.astype(np.float32))]
    y = (torch.rand(X.shape[0], X.shape[1], generator=torch.Generator().manual_seed(seed + 100)) < 0.3).float()
    net.train()
    net.zero_grad()
    out = net(inputs)
    loss = nn.BCELoss()(out, y)
    loss.backward()
    rec = {"X": X.astype(np.float32), "Ri": Ri.astype(np.uint8), "Ro": Ro.astype(np.uint8), "y": y.numpy(),
           "out": out.detach().numpy(), "loss": np.float32(loss.item()), "F": F, "h": h, "n_iters": n_iters, "seed": seed,
           "n_params": sum(p.numel() for p in net.parameters()), "keys": np.array(list(net.state_dict().keys()))}
    for k, v in net.state_dict().items():
        rec["param:" + k] = v.numpy()
    for k, v in net.named_parameters():
        rec["grad:" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-24s B=%d N=%d E=%d params=%d loss=%.6f out[0,:3]=%s" % (name, X.shape[0], X.shape[1], Ri.shape[2], rec["n_params"],
                                                                  loss.item(), out[0, :3].detach().numpy()))


def with_fourth_feature(graphs, seed):
    rng = np.random.RandomState(seed)
    out = []
    for g in graphs:
        X = np.concatenate([g.X, rng.uniform(-1, 1, (g.X.shape[0], 1)).astype(np.float32)], axis=1)
        out.append(g._replace(X=X))
    return out


def main():
    NC = reference_classes()
    os.makedirs(OUT, exist_ok=True)
    run_case(NC, "nodeclf_toy2d_h8_it1", data.toy2d_graphs(4, input_dim=3, seed=3), 3, 8, 1, 0)
    run_case(NC, "nodeclf_toy2d_h8_it0", data.toy2d_graphs(2, input_dim=3, seed=4), 3, 8, 0, 1)
    run_case(NC, "nodeclf_f4_h16_it2", with_fourth_feature(data.acts_like_graphs(2, n_tracks=12, seed=5), 6), 4, 16, 2, 2)
    run_case(NC, "nodeclf_acts_h32_it3", data.acts_like_graphs(3, n_tracks=30, seed=7), 3, 32, 3, 3)
    run_case(NC, "nodeclf_acts_h64_it2", data.acts_like_graphs(2, n_tracks=20, seed=9), 3, 64, 2, 4)


if __name__ == "__main__":
    main()
