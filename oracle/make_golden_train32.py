"""Generate tests/golden/train_grads_h32_it3.npz by RUNNING THE REFERENCE (TEST INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_train32.py      (build container: /root/reference exists)

One step of the reference's own Estimator.training_step (gnn/estimator.py:49-60: BCELoss over all padded slots + L1
penalty, loss.backward()) at hidden_dim 32 -- the width at which the CUDA backward runs its dense step on the tensor
cores -- on a ragged batch of three small ACTS-like events: inputs, parameters, the loss and the gradients.  Pins the
gradient oracle (oracle/segclf_oracle.sparse_vjp) at that width (tests/test_oracle_golden.py); the GPU tests compare
the CUDA backward with that oracle at hidden_dim 32 / 64 (tests/test_gpu_training.py).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import make_golden as G                         # noqa: E402  (the work-arounds and merge_graphs of the main generator)
import estimator as ref_estimator               # noqa: E402  (reference, on sys.path through make_golden)


def main():
    F, h, T = 3, 32, 3
    graphs = [G.data.acts_like_graph(n, seed=90 + i) for i, n in enumerate((20, 25, 16))]
    dense = [G.ref_graph.graph_from_sparse(G.ref_graph.SparseGraph(*g)) for g in graphs]
    X, Ri, Ro = G.merge_graphs(dense)
    e_max = Ri.shape[2]
    y = np.zeros((len(graphs), e_max), np.float32)
    for b, g in enumerate(graphs):
        y[b, :g.y.shape[0]] = g.y
    torch.manual_seed(13)
    ones = [torch.ones(h, 3 * (F + h)), torch.ones(h, h)]
    net = G.ref_model.SegmentClassifier(F, h, T, masks_e=None, masks_n=ones)
    rec = {"X": X.astype(np.float32), "Ri": Ri.astype(np.uint8), "Ro": Ro.astype(np.uint8), "y": y,
           "F": F, "h": h, "n_iters": T, "seed": 13, "l1": 1e-4}
    for k, v in net.state_dict().items():
        rec["param:" + k] = v.numpy().copy()
    with contextlib.redirect_stdout(io.StringIO()):
        est = ref_estimator.Estimator(net, torch.nn.BCELoss(), opt="Adam", l1=1e-4)
    inputs = [torch.from_numpy(a.astype(np.float32)) for a in (X, Ri, Ro)]
    net.train()
    loss = float(est.training_step(inputs, torch.from_numpy(y)).item())
    for k, p_ in net.named_parameters():
        rec["grad0:" + k] = p_.grad.numpy().copy()
    rec["losses"] = np.array([loss])
    np.savez_compressed(os.path.join(G.OUT, "train_grads_h32_it3.npz"), **rec)
    print("train_grads_h32_it3    loss=%r  nodes=%s edges=%d" % (loss, X.shape, e_max))


if __name__ == "__main__":
    main()
