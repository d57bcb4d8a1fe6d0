"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference exists:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

It imports the reference's own gnn/model.py, gnn/model_maskedlinear.py and gnn/graph.py
unmodified, builds seeded synthetic batches, and stores inputs, parameters and the
reference's outputs.  Nothing under tests/ or the package reads /root/reference at run time;
the committed .npz files are what travels to the GPU box.

Work-arounds needed to run the reference:
  * gnn/model.py:100 calls set_mask(None[0]) when masks_n is None -> pass all-ones masks_n
    (weight*1 == weight, bit-identical to the unmasked model);
  * model_maskedlinear has a module-global cuda=True -> set to False on this GPU-less host,
    and its prints are swallowed.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = "/root/reference/gnn"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import graph as ref_graph                      # noqa: E402  (reference)
import model as ref_model                      # noqa: E402  (reference)
import model_maskedlinear as ref_twin          # noqa: E402  (reference)

from gnn_fpga_b200 import data                 # noqa: E402  (input generators only)

ref_twin.cuda = False
torch.set_num_threads(1)   # fixed summation order inside the reference's bmm / addmm


def merge_graphs(graphs):
    """Same padding as gnn/trainSegmentClassifier.py:66-95 (that script cannot be imported:
    it needs graph.feature_scale, which gnn/graph.py:14-15 has commented out)."""
    B = len(graphs)
    if B == 1:
        g = graphs[0]
        return g.X[None], g.Ri[None], g.Ro[None]
    n_max = max(g.X.shape[0] for g in graphs)
    e_max = max(g.Ri.shape[1] for g in graphs)
    X = np.zeros((B, n_max, graphs[0].X.shape[1]), np.float32)
    Ri = np.zeros((B, n_max, e_max), np.uint8)
    Ro = np.zeros((B, n_max, e_max), np.uint8)
    for b, g in enumerate(graphs):
        n, e = g.Ri.shape
        X[b, :n], Ri[b, :n, :e], Ro[b, :n, :e] = g.X, g.Ri, g.Ro
    return X, Ri, Ro


def ones_masks(F, h):
    D = F + h
    return [torch.ones(h, 3 * D), torch.ones(h, h)]


def run_case(name, sparse_graphs, F, h, n_iters, seed, which="model", masks=None, extra=None):
    dense = [ref_graph.graph_from_sparse(ref_graph.SparseGraph(*g)) for g in sparse_graphs]
    X, Ri, Ro = merge_graphs(dense)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if which == "model":
            me, mn = (masks if masks is not None else (None, ones_masks(F, h)))
            net = ref_model.SegmentClassifier(F, h, n_iters, masks_e=me, masks_n=mn)
        else:
            me, mn = (masks if masks is not None else (None, None))
            net = ref_twin.SegmentClassifier(F, h, n_iters, masks_e=me, masks_n=mn)
        net.eval()
        with torch.no_grad():
            out = net([torch.from_numpy(X.astype(np.float32)), torch.from_numpy(Ri.astype(np.float32)),
                       torch.from_numpy(Ro.astype(np.float32))])
    n_params = sum(p.numel() for p in net.parameters())
    rec = {"X": X.astype(np.float32), "Ri": Ri.astype(np.uint8), "Ro": Ro.astype(np.uint8),
           "out": out.numpy(), "F": F, "h": h, "n_iters": n_iters, "seed": seed,
           "n_params": n_params, "which": which}
    for k, v in net.state_dict().items():
        rec["param:" + k] = v.numpy()
    if masks is not None:
        rec["mask_e0"], rec["mask_e1"] = masks[0][0].numpy(), masks[0][1].numpy()
        rec["mask_n0"], rec["mask_n1"] = masks[1][0].numpy(), masks[1][1].numpy()
    if extra:
        rec.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print("%-22s B=%d N=%d E=%d params=%d out[0,:3]=%s" % (name, X.shape[0], X.shape[1], Ri.shape[2], n_params, out[0, :3].numpy()))
    return net, out


def main():
    os.makedirs(OUT, exist_ok=True)
    # C1: BASELINE.json configs[0]
    toy3 = data.toy2d_graphs(32, input_dim=3, seed=0)
    run_case("c1_toy2d_h8_it1", toy3, 3, 8, 1, seed=0)
    # the Toy2D notebook's own configuration (F=2, h=32, n_iters=10)
    run_case("toy2d_f2_h32_it10", data.toy2d_graphs(4, input_dim=2, seed=1), 2, 32, 10, seed=1)
    # Inference.ipynb configuration (h=4, n_iters=1, 189 parameters)
    run_case("toy2d_h4_it1", data.toy2d_graphs(3, input_dim=3, seed=2), 3, 4, 1, seed=2)
    run_case("toy2d_h16_it3", data.toy2d_graphs(3, input_dim=3, seed=3), 3, 16, 3, seed=3, which="twin")
    # ragged ACTS-like batch -> exercises merge_graphs zero padding of nodes and edges
    ragged = [data.acts_like_graph(n, seed=10 + i) for i, n in enumerate((20, 31, 25))]
    run_case("acts_ragged_h32_it4", ragged, 3, 32, 4, seed=4, which="twin")
    # mu200 notebook configuration (h=64, n_iters=6) on small events
    run_case("acts_h64_it6", [data.acts_like_graph(30, seed=20), data.acts_like_graph(30, seed=21)], 3, 64, 6, seed=5)
    # masked twin (BASELINE.json configs[2]) with Bernoulli(0.5) masks, both reference files
    me, mn = data.random_masks(3, 8, keep=0.5, seed=1234)
    g_m = [data.acts_like_graph(24, seed=30 + i) for i in range(2)]
    _, o1 = run_case("acts_masked_h8_it4", g_m, 3, 8, 4, seed=6, which="model", masks=(me, mn))
    _, o2 = run_case("acts_masked_h8_it4_twin", g_m, 3, 8, 4, seed=6, which="twin", masks=(me, mn))
    assert torch.equal(o1, o2), "reference model.py and model_maskedlinear.py disagree"
    me, mn = data.random_masks(3, 32, keep=0.5, seed=99)
    run_case("acts_masked_h32_it4", g_m, 3, 32, 4, seed=7, which="model", masks=(me, mn))
    # half edges: a column of Ri (or Ro) that is entirely zero while the other is not
    g = data.acts_like_graph(12, seed=40)
    ri_r, ri_c, ro_r, ro_c = [np.array(a) for a in (g.Ri_rows, g.Ri_cols, g.Ro_rows, g.Ro_cols)]
    drop_in = ri_c != 5            # edge 5 loses its end
    drop_out = ro_c != 9           # edge 9 loses its start
    n_e = int(ri_c.max()) + 1
    half = ref_graph.SparseGraph(g.X, ri_r[drop_in], ri_c[drop_in], ro_r[drop_out], ro_c[drop_out], g.y)
    # graph_from_sparse sizes the matrices by len(Ri_rows) = n_e - 1 and would lose the last
    # column (IndexError), so the dense batch is filled by hand the way gnn/graph.py:132-135 does.
    Ri = np.zeros((g.X.shape[0], n_e), np.uint8)
    Ro = np.zeros((g.X.shape[0], n_e), np.uint8)
    Ri[half.Ri_rows, half.Ri_cols] = 1
    Ro[half.Ro_rows, half.Ro_cols] = 1
    torch.manual_seed(8)
    net = ref_model.SegmentClassifier(3, 8, 2, masks_e=None, masks_n=ones_masks(3, 8)).eval()
    with torch.no_grad():
        out = net([torch.from_numpy(g.X[None]), torch.from_numpy(Ri[None].astype(np.float32)),
                   torch.from_numpy(Ro[None].astype(np.float32))])
    rec = {"X": g.X[None], "Ri": Ri[None], "Ro": Ro[None], "out": out.numpy(), "F": 3, "h": 8, "n_iters": 2,
           "seed": 8, "n_params": sum(p.numel() for p in net.parameters()), "which": "model"}
    for k, v in net.state_dict().items():
        rec["param:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "half_edges_h8_it2.npz"), **rec)
    print("half_edges_h8_it2      out[0,4:10]=%s" % out[0, 4:10].numpy())

    # graph tuple round trip through the reference's graph.py (make_sparse_graph / graph_from_sparse
    # / save_graph / load_graph): the integer golden
    d = ref_graph.graph_from_sparse(ref_graph.SparseGraph(*ragged[1]))
    sg = ref_graph.make_sparse_graph(d.X, d.Ri, d.Ro, d.y)
    tmp = os.path.join(OUT, "ref_saved_graph.npz")
    ref_graph.save_graph((sg, None), tmp)     # written by the reference's own writer
    back = ref_graph.load_graph(tmp, ref_graph.SparseGraph)
    assert all(np.array_equal(a, b) for a, b in zip(sg, back))
    np.savez_compressed(os.path.join(OUT, "graph_roundtrip.npz"), X=d.X, Ri=d.Ri, Ro=d.Ro, y=d.y,
                        Ri_rows=sg.Ri_rows, Ri_cols=sg.Ri_cols, Ro_rows=sg.Ro_rows, Ro_cols=sg.Ro_cols)
    print("graph_roundtrip        N=%d E=%d dtype=%s" % (d.Ri.shape[0], d.Ri.shape[1], sg.Ri_rows.dtype))

    # two training steps through the reference's own Estimator.training_step (gnn/estimator.py:49-60):
    # BCELoss over all padded slots + L1 penalty, Adam
    import estimator as ref_estimator                                  # noqa: E402  (reference)
    g_t = [data.acts_like_graph(n, seed=50 + i) for i, n in enumerate((14, 18, 11))]
    dense = [ref_graph.graph_from_sparse(ref_graph.SparseGraph(*g)) for g in g_t]
    X, Ri, Ro = merge_graphs(dense)
    e_max = Ri.shape[2]
    y = np.zeros((len(g_t), e_max), np.float32)
    for b, g in enumerate(g_t):
        y[b, :g.y.shape[0]] = g.y
    torch.manual_seed(9)
    net = ref_model.SegmentClassifier(3, 8, 2, masks_e=None, masks_n=ones_masks(3, 8))
    rec = {"X": X.astype(np.float32), "Ri": Ri.astype(np.uint8), "Ro": Ro.astype(np.uint8), "y": y,
           "F": 3, "h": 8, "n_iters": 2, "seed": 9, "l1": 1e-4}
    for k, v in net.state_dict().items():
        rec["param:" + k] = v.numpy().copy()
    with contextlib.redirect_stdout(io.StringIO()):
        est = ref_estimator.Estimator(net, torch.nn.BCELoss(), opt="Adam", l1=1e-4)
    inputs = [torch.from_numpy(a.astype(np.float32)) for a in (X, Ri, Ro)]
    net.train()
    losses = []
    for step in range(2):
        losses.append(float(est.training_step(inputs, torch.from_numpy(y)).item()))
        if step == 0:
            for k, p_ in net.named_parameters():
                rec["grad0:" + k] = p_.grad.numpy().copy()
    for k, v in net.state_dict().items():
        rec["after:" + k] = v.numpy().copy()
    rec["losses"] = np.array(losses)
    np.savez_compressed(os.path.join(OUT, "train_step_h8_it2.npz"), **rec)
    print("train_step_h8_it2      losses=%s" % losses)

    # state_dict key list and parameter counts (SURVEY.md §4 structural pins)
    counts = {}
    for (F, h) in ((3, 4), (3, 8), (3, 32), (3, 64), (2, 32)):
        net = ref_model.SegmentClassifier(F, h, 1, masks_n=ones_masks(F, h))
        counts["%d,%d" % (F, h)] = sum(p.numel() for p in net.parameters())
    keys = list(net.state_dict().keys())
    np.savez(os.path.join(OUT, "structure.npz"), keys=np.array(keys), counts=np.array(list(counts.items())))
    print("structure", counts, keys)


if __name__ == "__main__":
    main()
