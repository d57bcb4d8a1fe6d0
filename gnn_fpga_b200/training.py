"""Training step for the drop-in SegmentClassifier, on the hand-written sm_100a kernels.

`Estimator.training_step` (gnn/estimator.py:49-60) does
    outputs = model(inputs); loss = loss_func(outputs, targets) + l1 * sum|W|; loss.backward();
    optimizer.step()
Two ways to run that here, both through the C ABI (include/gnnseg.h), neither with a CPU path:

* **drop-in**: `model(inputs)` under autograd returns a tensor whose `grad_fn` is `SegClfFunction`:
  forward = `gnnseg_forward_train` (the inference kernels, which also keep the per-iteration
  activations in the batch's training workspace), backward = `gnnseg_backward` (CSR-gather chain
  rule + per-CTA partial weight gradients summed in a fixed order: no atomics, bit-reproducible).
  Loss and optimiser stay whatever the caller's Estimator holds (torch's BCELoss / Adam).
* **fused**: `NativeTrainer.step(batch, targets)` also runs the loss (`gnnseg_bce_loss`,
  `gnnseg_l1_penalty`) and the optimiser (`gnnseg_adam_step` over one flat parameter buffer)
  through the library; the parameters of the module become views of that flat buffer so that one
  NCCL all-reduce and one Adam launch cover all ten tensors.

Multi-GPU: every rank holds its own shard of events; `allreduce_gradients`
averages the flattened gradients with one all-reduce (569 - 26 049 floats: latency bound).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .graph import _ptr, _stream_ptr


def _param_list(model):
    """The ten parameter tensors in GnnsegParams / state_dict order (gnn/model.py:128-138)."""
    e0, e2 = model.edge_network.network[0], model.edge_network.network[2]
    n0, n2 = model.node_network.network[0], model.node_network.network[2]
    lin = model.input_network[0]
    return [lin.weight, lin.bias, e0.weight, e0.bias, e2.weight, e2.bias, n0.weight, n0.bias, n2.weight, n2.bias]


def _mask_struct(model):
    """GnnsegParams with only the four mask pointers set (what gnnseg_backward reads)."""
    e0, e2 = model.edge_network.network[0], model.edge_network.network[2]
    n0, n2 = model.node_network.network[0], model.node_network.network[2]
    keep = [lin.effective_mask() for lin in (e0, e2, n0, n2)]
    ptrs = [None] * 10 + [m.data_ptr() if m is not None else None for m in keep]
    return _lib.GnnsegParams(*ptrs), keep


def _forward_train(model, batch):
    """gnnseg_forward_train on `batch`: returns (scores (n_slots,), blob, workspace)."""
    L = _lib.lib()
    dev = batch.device
    if batch.F != model.input_dim:
        raise ValueError("X has %d features, model expects input_dim=%d" % (batch.F, model.input_dim))
    blob = model.pack_weights()
    if blob.device != dev:
        raise ValueError("graph batch on %s but model on %s" % (dev, blob.device))
    h, T = model.hidden_dim, model.n_iters
    ws = batch.train_workspace(h, T)
    scores = torch.empty(batch.n_slots, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.gnnseg_forward_train(_ptr(blob), C.byref(batch.struct), _ptr(batch.X), batch.F, h, T,
                                          _ptr(scores), _ptr(ws), ws.numel(), _stream_ptr(dev)),
                   "gnnseg_forward_train")
    return scores, blob, ws


def _backward(model, batch, blob, ws, dscores, grads):
    """gnnseg_backward: `grads` = ten fp32 contiguous CUDA tensors shaped like the parameters."""
    L = _lib.lib()
    dev = batch.device
    masks, keep = _mask_struct(model)
    gs = _lib.GnnsegGrads(*[g.data_ptr() for g in grads])
    with torch.cuda.device(dev):
        _lib.check(L.gnnseg_backward(_ptr(blob), C.byref(masks), C.byref(batch.struct), batch.F, model.hidden_dim,
                                     model.n_iters, _ptr(dscores), C.byref(gs), _ptr(ws), ws.numel(),
                                     _stream_ptr(dev)), "gnnseg_backward")
    del keep


class SegClfFunction(torch.autograd.Function):
    """SegmentClassifier.forward (gnn/model.py:140-156) as one differentiable op."""

    @staticmethod
    def forward(ctx, model, batch, *params):
        scores, blob, ws = _forward_train(model, batch)
        ctx.model, ctx.batch, ctx.blob, ctx.ws = model, batch, blob, ws
        ctx.shapes = [p.shape for p in params]
        return scores.view(batch.B, batch.e_max)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        batch = ctx.batch
        dscores = grad_out.to(torch.float32).contiguous().view(-1)
        grads = [torch.empty(s, dtype=torch.float32, device=batch.device) for s in ctx.shapes]
        _backward(ctx.model, batch, ctx.blob, ctx.ws, dscores, grads)
        out = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[2:])]
        return (None, None, *out)


def differentiable_forward(model, batch):
    """What SegmentClassifier.forward runs when autograd needs the result (training mode)."""
    if batch.device.type != "cuda":
        raise _lib.GnnsegError("training runs on CUDA devices only: there is no CPU path")
    params = _param_list(model)
    for p in params:
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise ValueError("SegmentClassifier parameters must be contiguous fp32 tensors")
    return SegClfFunction.apply(model, batch, *params)


def l1_penalty(model):
    """gnn/estimator.py:54-56: sum |W| over every layer with a weight in the two networks."""
    ws = [l.weight for l in model.node_network.network if hasattr(l, "weight")]
    ws += [l.weight for l in model.edge_network.network if hasattr(l, "weight")]
    return sum(w.abs().sum() for w in ws)


def allreduce_gradients(params, group=None, n_local=None):
    """Average the gradients over the ranks with ONE all-reduce of a flat buffer.

    n_local = None: equal weights (every rank holds the same number of padded slots, the bench's weak
    scaling).  n_local = this rank's number of padded slots (B_local * E_max_local): the ranks' gradients
    are weighted n_local / sum(n_local), which is the gradient of the reference's loss -- nn.BCELoss() is a
    mean over ALL padded slots of the batch (gnn/estimator.py:57) -- over the union of the shards when the
    shards are uneven (`shard_bounds(weights=...)`).  The weight travels in the same all-reduce."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    if n_local is None:
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)
    else:
        w = torch.full((1,), float(n_local), dtype=grads[0].dtype, device=grads[0].device)
        flat = torch.cat([g.reshape(-1) * float(n_local) for g in grads] + [w])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat = flat[:-1] / flat[-1]
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def training_step(model, optimizer, loss_func, inputs, targets, l1=0.0, group=None, weight_by_slots=False):
    """Estimator.training_step (gnn/estimator.py:49-60) + gradient all-reduce when
    torch.distributed is initialised (every rank holds its own shard of events).  weight_by_slots:
    weight the ranks by their padded slot counts (uneven shards, see allreduce_gradients)."""
    model.zero_grad()
    optimizer.zero_grad()
    outputs = model(inputs)
    loss = loss_func(outputs, targets)
    if l1:
        loss = loss + l1 * l1_penalty(model)
    loss.backward()
    allreduce_gradients(list(model.parameters()), group, n_local=targets.numel() if weight_by_slots else None)
    optimizer.step()
    return loss


class NativeTrainer:
    """Estimator.training_step with every stage in the library: forward, BCE (mean over all padded
    slots) + L1 penalty, backward, gradient all-reduce, Adam.  No autograd graph is built.

    The module's ten parameters are re-pointed at views of one flat fp32 buffer (values kept), so
    the optimiser is one `gnnseg_adam_step` launch and the multi-GPU exchange one all-reduce.
    Hyper-parameters default to torch.optim.Adam's, which is what Estimator(opt='Adam') builds
    (gnn/estimator.py:35-36)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, l1=0.0, group=None,
                 weight_by_slots=False):
        dev = model._device()
        if dev.type != "cuda":
            raise _lib.GnnsegError("NativeTrainer needs the model on a CUDA device: there is no CPU path")
        self.model, self.group = model, group
        self.weight_by_slots = weight_by_slots       # uneven shards: weight the ranks by their padded slot counts
        self.lr, self.betas, self.eps, self.weight_decay, self.l1 = lr, betas, eps, weight_decay, float(l1)
        self.params = _param_list(model)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grads = []
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            self.grads.append(self.flat_grad[off:off + k].view(p.shape))
            off += k
        self.step_count = 0
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scratch = torch.empty(1024, dtype=torch.uint8, device=dev)
        self._dscores = None

    def _params_struct(self):
        m, keep = _mask_struct(self.model)
        ptrs = [p.data_ptr() for p in self.params] + [m.m_e1, m.m_e2, m.m_n1, m.m_n2]
        return _lib.GnnsegParams(*ptrs), keep

    def step(self, inputs, targets, weights=None):
        """One optimisation step on `inputs` (anything SegmentClassifier.forward accepts) against
        `targets` (B, E_max) fp32 on the device.  Returns the loss as a 1-element device tensor
        (no host synchronisation)."""
        L = _lib.lib()
        model = self.model
        batch = model._to_batch(inputs)
        dev = batch.device
        targets = targets.to(device=dev, dtype=torch.float32).contiguous()
        if targets.numel() != batch.n_slots:
            raise ValueError("targets has %d elements, the batch has %d slots" % (targets.numel(), batch.n_slots))
        if weights is not None:
            weights = weights.to(device=dev, dtype=torch.float32).contiguous()
        scores, blob, ws = _forward_train(model, batch)
        if self._dscores is None or self._dscores.numel() < batch.n_slots:
            self._dscores = torch.empty(batch.n_slots, dtype=torch.float32, device=dev)
        st = _stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_bce_loss(_ptr(scores), _ptr(targets), _ptr(weights) if weights is not None else None,
                                         batch.n_slots, _ptr(self.loss), _ptr(self._dscores), _ptr(self._scratch), st),
                       "gnnseg_bce_loss")
            _backward(model, batch, blob, ws, self._dscores, self.grads)
            if self.l1:
                ps, keep = self._params_struct()
                gs = _lib.GnnsegGrads(*[g.data_ptr() for g in self.grads])
                _lib.check(L.gnnseg_l1_penalty(C.byref(ps), model.input_dim, model.hidden_dim, self.l1, _ptr(self.loss),
                                               C.byref(gs), st), "gnnseg_l1_penalty")
                del keep
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                if self.weight_by_slots:
                    buf = torch.cat([self.flat_grad * float(batch.n_slots),
                                     torch.full((1,), float(batch.n_slots), dtype=torch.float32, device=dev)])
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                    self.flat_grad.copy_(buf[:-1] / buf[-1])
                else:
                    dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
                    self.flat_grad /= dist.get_world_size(self.group)
            self.step_count += 1
            _lib.check(L.gnnseg_adam_step(_ptr(self.flat), _ptr(self.flat_grad), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                          self.flat.numel(), self.step_count, self.lr, self.betas[0], self.betas[1],
                                          self.eps, self.weight_decay, st), "gnnseg_adam_step")
        self.scores = scores.view(batch.B, batch.e_max)
        return self.loss
