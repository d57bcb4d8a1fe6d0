"""Training step for the drop-in SegmentClassifier (interim, round 1).

`Estimator.training_step` (gnn/estimator.py:49-60) calls `model(inputs)` with autograd on, adds
an L1 penalty over the edge/node network weights, `loss.backward()`, `optimizer.step()`.  The
hand-written sm_100a kernels cover the forward only; until their backward twins exist
 a forward under autograd runs `autograd_forward` below: the SAME sparse
formulation (int32 endpoints produced by the CUDA graph kernels, gathers + ordered index_add)
written with torch ops ON THE GPU, so that autograd can differentiate it.  It never touches the
dense (B,N,E) incidence tensors beyond the one conversion, never runs on the CPU, and is not
used for inference.  Multi-GPU: `allreduce_gradients` averages the flattened gradients with
one NCCL all-reduce (569 - 26 049 floats: latency bound), as SURVEY.md §8(e) specifies.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F


def _eff(lin):
    """MaskedLinear's effective weight (gnn/model.py:28-31)."""
    m = lin.effective_mask() if hasattr(lin, "effective_mask") else None
    return lin.weight * m if m is not None else lin.weight


def _rows(H, idx):
    """H[idx] with the zero row for idx == -1 (an absent endpoint = zero column of Ri/Ro)."""
    return H[idx.clamp(min=0)] * (idx >= 0).to(H.dtype).unsqueeze(1)


def autograd_forward(model, batch):
    """gnn/model.py:140-156 on a DeviceGraphBatch with torch ops (differentiable).  Returns
    (B, e_max) scores including the padded slots, like the reference on merge_graphs' batches."""
    if batch.device.type != "cuda":
        raise RuntimeError("training runs on CUDA devices only")
    src, dst = batch.src.long(), batch.dst.long()
    X = batch.X
    e0, e2 = model.edge_network.network[0], model.edge_network.network[2]
    n0, n2 = model.node_network.network[0], model.node_network.network[2]
    lin_in = model.input_network[0]
    real_i, real_o = dst >= 0, src >= 0

    def edge(H):
        B_ = torch.cat([_rows(H, src), _rows(H, dst)], dim=1)           # [bo | bi], gnn/model.py:71-73
        z = torch.tanh(F.linear(B_, _eff(e0), e0.bias))
        return torch.sigmoid(F.linear(z, _eff(e2), e2.bias)).squeeze(-1)

    def node(H, e):
        w = e.unsqueeze(1)
        mi = torch.zeros_like(H).index_add(0, dst[real_i], (w * _rows(H, src))[real_i])   # :118
        mo = torch.zeros_like(H).index_add(0, src[real_o], (w * _rows(H, dst))[real_o])   # :119
        z = torch.tanh(F.linear(torch.cat([mi, mo, H], dim=1), _eff(n0), n0.bias))
        return torch.tanh(F.linear(z, _eff(n2), n2.bias))

    H = torch.cat([torch.tanh(F.linear(X, lin_in.weight, lin_in.bias)), X], dim=1)
    for _ in range(model.n_iters):
        e = edge(H)
        H = torch.cat([node(H, e), X], dim=1)
    return edge(H).view(batch.B, batch.e_max)


def l1_penalty(model):
    """gnn/estimator.py:54-56: sum |W| over every layer with a weight in the two networks."""
    ws = [l.weight for l in model.node_network.network if hasattr(l, "weight")]
    ws += [l.weight for l in model.edge_network.network if hasattr(l, "weight")]
    return sum(w.abs().sum() for w in ws)


def allreduce_gradients(params, group=None):
    """Average the gradients over the ranks with ONE all-reduce of a flat buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def training_step(model, optimizer, loss_func, inputs, targets, l1=0.0, group=None):
    """Estimator.training_step (gnn/estimator.py:49-60) + gradient all-reduce when
    torch.distributed is initialised (every rank holds its own shard of events)."""
    model.zero_grad()
    optimizer.zero_grad()
    outputs = model(inputs)
    loss = loss_func(outputs, targets)
    if l1:
        loss = loss + l1 * l1_penalty(model)
    loss.backward()
    allreduce_gradients(list(model.parameters()), group)
    optimizer.step()
    return loss
