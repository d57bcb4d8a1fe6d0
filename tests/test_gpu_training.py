"""Parity of the native training path (gnnseg_forward_train / gnnseg_backward / gnnseg_bce_loss /
gnnseg_l1_penalty / gnnseg_adam_step, all through the C ABI) with the reference's Estimator
(tests/golden/train_step_h8_it2.npz) and with autograd over the oracle's sparse restatement.

Tolerances: gradients <= 2e-5 of the tensor's largest entry against the fp64 oracle (fp32 sums over
up to ~10^5 nodes); losses 1e-5 relative; bit-equal from run to run.  Everything here needs a GPU.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case
from oracle import segclf_oracle as O

pytestmark = pytest.mark.gpu
GTOL = 2e-5


def grad_err(a, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(np.asarray(a, np.float64) - ref)) / (np.max(np.abs(ref)) + 1e-30))


def _model(F, h, T, params, device, masks_e=None, masks_n=None):
    from gnn_fpga_b200 import SegmentClassifier
    m = SegmentClassifier(F, h, T, masks_e=masks_e, masks_n=masks_n)
    m.load_state_dict(params)
    return m.to(device).train()


CASES = [
    # F, h, T, tracks per event, masked, half edges
    (3, 8, 2, (14, 18, 11), False, True),
    (3, 32, 4, (60, 45), False, False),
    (3, 64, 2, (50,), False, False),
    (2, 4, 1, (20, 9), False, False),
    (3, 16, 3, (30, 33), True, False),
    (3, 8, 0, (12, 12), False, False),
    (4, 32, 1, (25, 40), True, True),
]


@pytest.mark.parametrize("F,h,T,tracks,masked,half", CASES)
def test_backward_matches_autograd_oracle(F, h, T, tracks, masked, half, cuda_device):
    """Random cotangent through SegClfFunction against torch autograd (CPU, fp64) over the sparse
    restatement: ragged batches (padding slots), half edges (one endpoint absent), masks."""
    from gnn_fpga_b200 import data
    graphs = [data.acts_like_graph(n, seed=70 + i) for i, n in enumerate(tracks)]
    if F != 3:
        graphs = [g._replace(X=np.ascontiguousarray(np.concatenate([g.X, g.X[:, :1] * 0.5], 1)[:, :F])) for g in graphs]
    if half:   # drop the start of one edge and the end of another (a zero column of Ro / Ri); graphs[1] sets e_max
        g = graphs[0]
        graphs[0] = g._replace(Ro_rows=g.Ro_rows[g.Ro_cols != 3], Ro_cols=g.Ro_cols[g.Ro_cols != 3],
                               Ri_rows=g.Ri_rows[g.Ri_cols != 5], Ri_cols=g.Ri_cols[g.Ri_cols != 5])
    p = O.init_params(F, h, seed=3)
    masks_e = masks_n = None
    if masked:
        masks_e, masks_n = data.random_masks(F, h, keep=0.6, seed=5)
    model = _model(F, h, T, p, cuda_device, masks_e, masks_n)
    X, src, dst, e_max = O.flatten_sparse_batch(graphs)
    rng = np.random.RandomState(1)
    cot = rng.normal(size=src.shape[0]).astype(np.float32)
    out = model(graphs)
    assert out.requires_grad and out.shape == (len(graphs), e_max)
    out.backward(torch.from_numpy(cot).to(cuda_device).view_as(out))
    ref_out, _, ref = O.sparse_vjp(p, X, src, dst, T, dscores=cot, masks_e=masks_e, masks_n=masks_n)
    assert np.max(np.abs(out.detach().cpu().numpy().reshape(-1) - ref_out.numpy()) / np.abs(ref_out.numpy())) <= 1e-5
    for k, v in model.named_parameters():
        assert v.grad is not None and v.grad.shape == v.shape, k
        assert grad_err(v.grad.cpu().numpy(), ref[k].numpy()) <= GTOL, (k, grad_err(v.grad.cpu().numpy(), ref[k].numpy()))
    if masked:     # masked-out weights get exactly zero gradient
        g = model.edge_network.network[0].weight.grad
        assert torch.all(g[masks_e[0].to(cuda_device) == 0] == 0)


def _golden_training_setup(cuda_device):
    z = np.load(os.path.join(GOLDEN, "train_step_h8_it2.npz"))
    params = {k[6:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("param:")}
    model = _model(int(z["F"]), int(z["h"]), int(z["n_iters"]), params, cuda_device)
    inputs = [torch.from_numpy(z[k].astype(np.float32)).to(cuda_device) for k in ("X", "Ri", "Ro")]
    y = torch.from_numpy(z["y"]).to(cuda_device)
    return z, model, inputs, y


def test_training_step_matches_reference_estimator(cuda_device):
    """Two optimisation steps as Estimator.training_step does them (gnn/estimator.py:49-60: torch's
    BCELoss over all padded slots + L1 penalty, torch's Adam) with the model's forward / backward
    on the CUDA kernels, against losses, first gradients and final parameters recorded from the
    reference's own Estimator."""
    from gnn_fpga_b200.training import training_step
    z, model, inputs, y = _golden_training_setup(cuda_device)
    opt = torch.optim.Adam(model.parameters())
    losses = []
    for step in range(2):
        loss = training_step(model, opt, torch.nn.BCELoss(), inputs, y, l1=float(z["l1"]))
        losses.append(float(loss.item()))
        if step == 0:
            for k, p_ in model.named_parameters():
                assert grad_err(p_.grad.cpu().numpy(), z["grad0:" + k]) <= GTOL, k
    assert np.allclose(losses, z["losses"], rtol=1e-5)
    for k, v in model.state_dict().items():
        assert np.allclose(v.cpu().numpy(), z["after:" + k], rtol=1e-4, atol=1e-6), k
    # inference after training: eval() goes through gnnseg_forward and agrees with the training forward
    model.eval()
    with torch.no_grad():
        a = model(inputs)
    model.train()
    b = model(inputs).detach()
    assert float(((a - b).abs() / b.abs()).max()) <= 1e-6


def test_native_trainer_matches_reference_estimator(cuda_device):
    """The fused step (library BCE + L1 + Adam, no autograd) against the same recording."""
    from gnn_fpga_b200.training import NativeTrainer
    z, model, inputs, y = _golden_training_setup(cuda_device)
    tr = NativeTrainer(model, l1=float(z["l1"]))
    losses = []
    for step in range(2):
        losses.append(float(tr.step(inputs, y).item()))
        if step == 0:
            for (k, _), g in zip(model.named_parameters(), tr.grads):
                assert grad_err(g.cpu().numpy(), z["grad0:" + k]) <= GTOL, k
    assert np.allclose(losses, z["losses"], rtol=1e-5)
    for k, v in model.state_dict().items():
        assert np.allclose(v.cpu().numpy(), z["after:" + k], rtol=1e-4, atol=1e-6), k


def test_backward_is_deterministic(cuda_device):
    """No atomics, fixed-order partial sums: two runs give bit-identical gradients."""
    from gnn_fpga_b200 import DeviceGraphBatch, data
    graphs = data.acts_like_graphs(6, n_tracks=150, seed=3)
    p = O.init_params(3, 32, seed=1)
    model = _model(3, 32, 3, p, cuda_device)
    batch = DeviceGraphBatch.from_sparse_graphs(graphs, cuda_device)
    cot = torch.randn(batch.B, batch.e_max, device=cuda_device, generator=torch.Generator(cuda_device).manual_seed(0))
    runs = []
    for _ in range(2):
        model.zero_grad()
        model(batch).backward(cot)
        runs.append([q.grad.clone() for q in model.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)


def test_event_scale_gradients(cuda_device):
    """One 4k-hit ACTS-like event, h=32, n_iters=4 (BASELINE configs[1] shape): loss and gradients of
    the fused step's first half against the fp64 oracle."""
    from gnn_fpga_b200 import data
    g = data.acts_like_graph(400, seed=11)
    p = O.init_params(3, 32, seed=0)
    model = _model(3, 32, 4, p, cuda_device)
    X, src, dst, e_max = O.flatten_sparse_batch([g])
    y = g.y.reshape(1, -1)
    out = model([g])
    loss = torch.nn.functional.binary_cross_entropy(out, torch.from_numpy(y).to(cuda_device))
    loss.backward()
    _, ref_loss, ref = O.sparse_vjp(p, X, src, dst, 4, y=y)
    assert abs(float(loss.item()) - float(ref_loss)) <= 1e-5 * float(ref_loss)
    for k, v in model.named_parameters():
        assert grad_err(v.grad.cpu().numpy(), ref[k].numpy()) <= 5e-5, (k, grad_err(v.grad.cpu().numpy(), ref[k].numpy()))


def test_loss_and_optimizer_kernels_against_torch(cuda_device):
    """gnnseg_bce_loss (value + gradient, clamped logs, optional weights) and gnnseg_adam_step
    against torch.nn.functional.binary_cross_entropy / torch.optim.Adam on the same numbers."""
    import ctypes as C
    from gnn_fpga_b200 import _lib
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    L = _lib.lib()
    gen = torch.Generator(cuda_device).manual_seed(4)
    n = 70001
    p = torch.rand(n, device=cuda_device, generator=gen).clamp(1e-6, 1 - 1e-6)
    p[:3] = torch.tensor([0.0, 1.0, 0.5], device=cuda_device)            # saturated scores hit the clamps
    y = (torch.rand(n, device=cuda_device, generator=gen) < 0.3).float()
    w = torch.rand(n, device=cuda_device, generator=gen)
    loss = torch.zeros(1, device=cuda_device)
    dp = torch.empty(n, device=cuda_device)
    ws = torch.empty(1024, dtype=torch.uint8, device=cuda_device)
    for weights in (None, w):
        assert L.gnnseg_bce_loss(_ptr(p), _ptr(y), _ptr(weights) if weights is not None else None, n, _ptr(loss),
                                 _ptr(dp), _ptr(ws), _stream_ptr(cuda_device)) == 0
        pr = p.clone().requires_grad_(True)
        ref = torch.nn.functional.binary_cross_entropy(pr, y, weight=weights)
        ref.backward()
        assert abs(float(loss.item()) - float(ref.item())) <= 2e-6 * float(ref.item())
        assert torch.allclose(dp, pr.grad, rtol=1e-5, atol=1e-12)
    # Adam, three steps, with and without weight decay
    for wd in (0.0, 0.01):
        x0 = torch.randn(5000, device=cuda_device, generator=gen)
        ref_p = x0.clone().requires_grad_(True)
        opt = torch.optim.Adam([ref_p], lr=2e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=wd)
        mine, m, v = x0.clone(), torch.zeros_like(x0), torch.zeros_like(x0)
        for step in range(1, 4):
            g = torch.randn(5000, device=cuda_device, generator=gen)
            ref_p.grad = g.clone()
            opt.step()
            assert L.gnnseg_adam_step(_ptr(mine), _ptr(g), _ptr(m), _ptr(v), 5000, step, 2e-3, 0.8, 0.95, 1e-7, wd,
                                      _stream_ptr(cuda_device)) == 0
        assert torch.allclose(mine, ref_p.detach(), rtol=1e-5, atol=1e-7)


def test_l1_penalty_kernel(cuda_device):
    import ctypes as C
    from gnn_fpga_b200 import _lib
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    from gnn_fpga_b200.training import _param_list
    L = _lib.lib()
    model = _model(3, 16, 1, O.init_params(3, 16, seed=2), cuda_device)
    params = _param_list(model)
    with torch.no_grad():
        params[2][0, 0] = 0.0                                          # sign(0) = 0
    ps = _lib.GnnsegParams(*([q.data_ptr() for q in params] + [None] * 4))
    grads = [torch.ones_like(q) for q in params]
    gs = _lib.GnnsegGrads(*[g.data_ptr() for g in grads])
    loss = torch.full((1,), 2.0, device=cuda_device)
    assert L.gnnseg_l1_penalty(C.byref(ps), 3, 16, 0.25, _ptr(loss), C.byref(gs), _stream_ptr(cuda_device)) == 0
    ws = [params[i] for i in (2, 4, 6, 8)]
    ref = 2.0 + 0.25 * sum(float(q.detach().abs().sum()) for q in ws)
    assert abs(float(loss.item()) - ref) <= 1e-5 * ref
    for i in (2, 4, 6, 8):
        assert torch.equal(grads[i], 1.0 + 0.25 * torch.sign(params[i].detach()))
    for i in (0, 1, 3, 5, 7, 9):
        assert torch.equal(grads[i], torch.ones_like(grads[i]))


_DENSE_AB_SCRIPT = """
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, data
from oracle import segclf_oracle as O
dev = torch.device("cuda:0")
out = {{}}
for h, T, n_ev in ((32, 4, 12), (64, 2, 3)):
    graphs = data.acts_like_graphs(n_ev, n_tracks=400, seed=21)
    m = SegmentClassifier(3, h, T)
    m.load_state_dict(O.init_params(3, h, seed=2))
    m = m.to(dev).train()
    batch = DeviceGraphBatch.from_sparse_graphs(graphs, dev)
    cot = torch.randn(batch.B, batch.e_max, device=dev, generator=torch.Generator(dev).manual_seed(5))
    for rep in range(2):
        m.zero_grad()
        m(batch).backward(cot)
        for k, v in m.named_parameters():
            out["h%d:%d:%s" % (h, rep, k)] = v.grad.cpu().numpy()
np.savez({dst!r}, **out)
"""


def test_dense_backward_tensor_core_paths_against_simt(tmp_path, cuda_device):
    """The dense backward step on tcgen05 (gnnseg_dprop_tc.cu + gnnseg_wgrad_tc.cu, the default at hidden_dim 32; weight
    gradients only at 64) against the all-SIMT fp32 kernel (GNNSEG_DENSE_BWD=simt) and the mixed form (=w) at batch scale:
    12 events of 4 000 hits (48 000 nodes: every CTA sees several tiles and the partial last stage), 3xTF32 against fp32
    within 2e-5 of each tensor's largest entry, and every path bit-reproducible from run to run.  The switch is read
    once per process: three child processes."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for mode in ("tc", "w", "simt"):
        dst = str(tmp_path / ("grads_%s.npz" % mode))
        env = dict(os.environ, GNNSEG_DENSE_BWD=mode)
        subprocess.run([sys.executable, "-c", _DENSE_AB_SCRIPT.format(root=root, dst=dst)], check=True, env=env, timeout=300)
        res[mode] = np.load(dst)
    keys = [k for k in res["simt"].files if ":0:" in k]
    assert len(keys) == 20
    for k in keys:
        for mode in ("tc", "w", "simt"):
            assert np.array_equal(res[mode][k], res[mode][k.replace(":0:", ":1:")]), (mode, k)
        for mode in ("tc", "w"):
            assert grad_err(res[mode][k], res["simt"][k]) <= GTOL, (mode, k, grad_err(res[mode][k], res["simt"][k]))
