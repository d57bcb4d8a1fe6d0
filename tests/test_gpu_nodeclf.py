"""Parity of the NodeClassifier path (gnnseg_pack_node_head / gnnseg_forward_nodes /
gnnseg_forward_nodes_train / gnnseg_backward_nodes, through the C ABI) with outputs, loss and
gradients of the reference's own NodeClassifier (gnn/MPNN_HitClassifier.ipynb c21, recorded in
tests/golden/nodeclf_*.npz) and with the oracle's sparse restatement on ragged batches.

Tolerances: node scores 1e-5 relative; gradients 2e-5 of each tensor's largest entry.
"""
import numpy as np
import pytest
import torch

from conftest import NODECLF_CASES, load_case, rel_err
from oracle import segclf_oracle as O

pytestmark = pytest.mark.gpu
GTOL = 2e-5


def grad_err(a, ref, scale_like=None):
    ref = np.asarray(ref, np.float64)
    scale = np.max(np.abs(ref))
    if scale_like is not None:      # a bias: see _scale_tensor
        scale = max(scale, float(np.max(np.abs(scale_like))))
    return float(np.max(np.abs(np.asarray(a, np.float64) - ref)) / (scale + 1e-30))


def _scale_tensor(grads, key):
    """A bias gradient is the plain sum of the terms whose activation-weighted sum (|act| <= 1) is
    the layer's weight gradient; when the sum cancels (a 1-element bias can end up 100x below its
    terms) the fp32 error is still that of the terms, so a bias is also judged on the scale of its
    layer's weight gradient."""
    return grads.get(key[:-len("bias")] + "weight") if key.endswith("bias") else None


def _model(rec_or_params, F, h, T, device):
    from gnn_fpga_b200.node_classifier import NodeClassifier
    m = NodeClassifier(F, h, T)
    m.load_state_dict(rec_or_params)
    return m.to(device)


@pytest.mark.parametrize("name", NODECLF_CASES)
def test_forward_matches_reference(name, cuda_device):
    rec = load_case(name)
    model = _model(rec["params"], rec["F"], rec["h"], rec["n_iters"], cuda_device).eval()
    assert list(model.state_dict().keys()) == list(rec["keys"])
    assert sum(p.numel() for p in model.parameters()) == rec["n_params"]
    inputs = [torch.from_numpy(rec[k].astype(np.float32)).to(cuda_device) for k in ("X", "Ri", "Ro")]
    with torch.no_grad():
        out = model(inputs)
    assert out.shape == rec["out"].shape and out.dtype == torch.float32
    assert rel_err(out.cpu().numpy(), rec["out"]) <= 1e-5


@pytest.mark.parametrize("name", NODECLF_CASES)
def test_training_backward_matches_reference(name, cuda_device):
    """model(inputs) under autograd + torch's BCELoss, as the notebook's Estimator runs it."""
    rec = load_case(name)
    model = _model(rec["params"], rec["F"], rec["h"], rec["n_iters"], cuda_device).train()
    inputs = [torch.from_numpy(rec[k].astype(np.float32)).to(cuda_device) for k in ("X", "Ri", "Ro")]
    out = model(inputs)
    assert out.requires_grad
    loss = torch.nn.BCELoss()(out, torch.from_numpy(rec["y"]).to(cuda_device))
    loss.backward()
    assert rel_err(out.detach().cpu().numpy(), rec["out"]) <= 1e-5
    assert abs(loss.item() - float(rec["loss"])) <= 1e-5 * abs(float(rec["loss"]))
    for k, v in model.named_parameters():
        ref = rec["grads"][k]
        assert v.grad is not None and v.grad.shape == ref.shape, k
        if np.max(np.abs(ref)) == 0:       # n_iters = 0: the edge / node networks are unused
            assert float(v.grad.abs().max()) == 0, k
        else:
            err = grad_err(v.grad.cpu().numpy(), ref, _scale_tensor(rec["grads"], k))
            assert err <= GTOL, (k, err)


@pytest.mark.parametrize("F,h,T,tracks", [(3, 32, 2, (40, 25, 33)), (3, 64, 1, (30, 45)), (4, 8, 3, (20, 9)), (3, 16, 2, (700,))])
def test_ragged_sparse_batches_against_oracle(F, h, T, tracks, cuda_device):
    """Host SparseGraph tuples with different node counts: flat per-node scores, random cotangent."""
    from gnn_fpga_b200 import data
    graphs = [data.acts_like_graph(n, seed=40 + i) for i, n in enumerate(tracks)]
    if F != 3:
        graphs = [g._replace(X=np.ascontiguousarray(np.concatenate([g.X, g.X[:, :1] * 0.5], 1)[:, :F])) for g in graphs]
    torch.manual_seed(11)
    p = O.init_params(F, h, seed=5)
    head = torch.nn.Linear(F + h, 1)
    p[O.HEAD_KEYS[0]], p[O.HEAD_KEYS[1]] = head.weight.detach().clone(), head.bias.detach().clone()
    model = _model(p, F, h, T, cuda_device).train()
    X, src, dst, _ = O.flatten_sparse_batch(graphs)
    cot = np.random.RandomState(2).normal(size=X.shape[0]).astype(np.float32)
    out = model(graphs)
    n_total = sum(10 * n for n in tracks)
    assert out.shape == ((1, n_total) if len(tracks) == 1 else (n_total,))
    out.backward(torch.from_numpy(cot).to(cuda_device).view_as(out))
    ref_out, _, ref = O.nodeclf_sparse_vjp(p, X, src, dst, T, dnode=cot)
    assert rel_err(out.detach().cpu().numpy().reshape(-1), ref_out.numpy()) <= 1e-5
    refn = {k: v.numpy() for k, v in ref.items()}
    for k, v in model.named_parameters():
        err = grad_err(v.grad.cpu().numpy(), refn[k], _scale_tensor(refn, k))
        assert err <= GTOL, (k, err)
    # inference path (gnnseg_forward_nodes) gives the training forward's scores bit for bit
    with torch.no_grad():
        again = model.eval()(graphs)
    assert torch.equal(again, out.detach())


def test_backward_is_deterministic(cuda_device):
    from gnn_fpga_b200 import data
    graphs = [data.acts_like_graph(60, seed=3), data.acts_like_graph(60, seed=4)]
    rec = load_case("nodeclf_acts_h32_it3")
    model = _model(rec["params"], 3, 32, 3, cuda_device).train()
    runs = []
    for _ in range(2):
        model.zero_grad()
        out = model(graphs)
        out.sum().backward()
        runs.append([v.grad.clone() for v in model.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)


def test_events_without_edges(cuda_device):
    """n_slots = 0 with n_iters > 0: the node steps see no messages, the backward has no edge launches."""
    from gnn_fpga_b200 import SparseGraph
    rng = np.random.RandomState(5)
    e = np.zeros(0, np.int64)
    graphs = [SparseGraph(rng.uniform(-1, 1, (n, 3)).astype(np.float32), e, e, e, e, np.zeros(0, np.float32)) for n in (37, 37)]
    torch.manual_seed(3)
    p = O.init_params(3, 32, seed=8)
    head = torch.nn.Linear(35, 1)
    p[O.HEAD_KEYS[0]], p[O.HEAD_KEYS[1]] = head.weight.detach().clone(), head.bias.detach().clone()
    model = _model(p, 3, 32, 2, cuda_device).train()
    out = model(graphs)
    assert out.shape == (2, 37)
    cot = rng.normal(size=74).astype(np.float32)
    out.backward(torch.from_numpy(cot).to(cuda_device).view_as(out))
    X = np.concatenate([g.X for g in graphs])
    ref_out, _, ref = O.nodeclf_sparse_vjp(p, X, e, e, 2, dnode=cot)
    assert rel_err(out.detach().cpu().numpy().reshape(-1), ref_out.numpy()) <= 1e-5
    refn = {k: v.numpy() for k, v in ref.items()}
    for k, v in model.named_parameters():
        if np.max(np.abs(refn[k])) == 0:
            assert float(v.grad.abs().max()) == 0, k
        else:
            err = grad_err(v.grad.cpu().numpy(), refn[k], _scale_tensor(refn, k))
            assert err <= GTOL, (k, err)
