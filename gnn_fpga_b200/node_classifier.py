"""Drop-in NodeClassifier (gnn/MPNN_HitClassifier.ipynb cell 21): binary classification of hits.

The model is the segment classifier's body -- input network, then `n_iters` x (edge network, node
network), every hidden state re-concatenated with X (cells 20-21 of the notebook are the same
EdgeNetwork / NodeNetwork as gnn/model.py:36-125, without masks) -- followed by
`output_network = Sequential(Linear(input_dim + hidden_dim, 1), Sigmoid())` applied to every node,
and NO final edge step.  Constructor signature, module tree, parameter order and `state_dict` keys
are the notebook's, so its checkpoints load.

Execution: the same CUDA kernels as SegmentClassifier (include/gnnseg.h).  The head's logit is one
more linear map of [H | X]; `gnnseg_pack_node_head` packs it into column 0 of the last step's
start-node projection, `gnnseg_forward_nodes` runs the steps and applies the sigmoid.  Under
autograd the forward is `gnnseg_forward_nodes_train` and the result's grad_fn runs
`gnnseg_backward_nodes`.  There is no CPU path.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .graph import DeviceGraphBatch, _ptr, _stream_ptr
from .model import SegmentClassifier
from .training import _mask_struct, _param_list


class NodeClassifier(SegmentClassifier):
    """`forward(inputs)` takes what SegmentClassifier.forward takes; it returns one score per node:
    (B, N) for `[X, Ri, Ro]` and for batches whose events all have N nodes (the reference call),
    otherwise a flat (n_nodes,) tensor in event order (`batch.n_nodes_per_event` splits it)."""

    def __init__(self, input_dim=4, hidden_dim=8, n_iters=1, hidden_activation=nn.Tanh):
        super().__init__(input_dim, hidden_dim, n_iters, hidden_activation)
        self.output_network = nn.Sequential(nn.Linear(input_dim + hidden_dim, 1), nn.Sigmoid())
        self._head_blob = None

    def pack_head(self):
        """gnnseg_pack_node_head: the blob of the last producing step (head in P column 0)."""
        L = _lib.lib()
        dev = self._device()
        n = L.gnnseg_weights_floats(self.input_dim, self.hidden_dim)
        if self._head_blob is None or self._head_blob.device != dev or self._head_blob.numel() != n:
            self._head_blob = torch.empty(n, dtype=torch.float32, device=dev)
        params, keep = self._params_struct()
        w, b = self.output_network[0].weight.detach(), self.output_network[0].bias.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.to(torch.float32).contiguous()
        if b.dtype != torch.float32 or not b.is_contiguous():
            b = b.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_pack_node_head(C.byref(params), _ptr(w), _ptr(b), self.input_dim, self.hidden_dim,
                                               _ptr(self._head_blob), _stream_ptr(dev)), "gnnseg_pack_node_head")
        del keep
        return self._head_blob

    @staticmethod
    def _shape(batch, scores):
        per = batch.n_nodes_per_event
        if per is not None and len(per) == batch.B and len(set(int(v) for v in per)) == 1:
            return scores.view(batch.B, int(per[0]))
        return scores

    def _run(self, batch):
        L = _lib.lib()
        dev = batch.device
        if batch.F != self.input_dim:
            raise ValueError("X has %d features, model expects input_dim=%d" % (batch.F, self.input_dim))
        blob = self.pack_weights()
        head = self.pack_head()
        if blob.device != dev:
            raise ValueError("graph batch on %s but model on %s" % (dev, blob.device))
        h = self.hidden_dim
        ws = batch.workspace(h)
        out = torch.empty(batch.n_nodes, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_forward_nodes(_ptr(blob), _ptr(head), C.byref(batch.struct), _ptr(batch.X), batch.F, h,
                                              self.n_iters, _ptr(out), _ptr(ws), ws.numel(), _stream_ptr(dev)),
                       "gnnseg_forward_nodes")
        return _original_order(batch, out)

    def forward(self, inputs):
        if self._device().type != "cuda":
            raise _lib.GnnsegError("NodeClassifier parameters are on %s: there is no CPU path" % self._device())
        if _lib.lib().gnnseg_supported(self.input_dim, self.hidden_dim) == 0:
            _lib.check(-2, "NodeClassifier(input_dim=%d, hidden_dim=%d)" % (self.input_dim, self.hidden_dim))
        batch = self._to_batch(inputs)
        if self._wants_grad():
            params = _param_list(self) + [self.output_network[0].weight, self.output_network[0].bias]
            for p in params:
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("NodeClassifier parameters must be contiguous fp32 tensors")
            return self._shape(batch, NodeClfFunction.apply(self, batch, *params))
        return self._shape(batch, self._run(batch))

    def predict_stream(self, batches, depth=2):
        raise NotImplementedError("predict_stream is the segment classifier's pipeline; call the model per batch")


def _original_order(batch, per_node):
    """Per-node values of a batch whose nodes were renumbered internally (GraphStore reorder) -> the
    order of the caller's X rows."""
    order = batch.original_node_ids()
    if order is None:
        return per_node
    out = torch.empty_like(per_node)
    out[order] = per_node
    return out


class NodeClfFunction(torch.autograd.Function):
    """NodeClassifier.forward as one differentiable op (gnnseg_forward_nodes_train / gnnseg_backward_nodes)."""

    @staticmethod
    def forward(ctx, model, batch, *params):
        L = _lib.lib()
        dev = batch.device
        if batch.F != model.input_dim:
            raise ValueError("X has %d features, model expects input_dim=%d" % (batch.F, model.input_dim))
        blob, head = model.pack_weights(), model.pack_head()
        h, T = model.hidden_dim, model.n_iters
        ws = batch.train_workspace(h, T)
        out = torch.empty(batch.n_nodes, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_forward_nodes_train(_ptr(blob), _ptr(head), C.byref(batch.struct), _ptr(batch.X), batch.F,
                                                    h, T, _ptr(out), _ptr(ws), ws.numel(), _stream_ptr(dev)),
                       "gnnseg_forward_nodes_train")
        ctx.model, ctx.batch, ctx.blob, ctx.head, ctx.ws = model, batch, blob, head, ws
        ctx.shapes = [p.shape for p in params]
        return _original_order(batch, out)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        L = _lib.lib()
        model, batch = ctx.model, ctx.batch
        dev = batch.device
        dnode = grad_out.to(torch.float32).contiguous().view(-1)
        if batch.original_node_ids() is not None:
            dnode = dnode[batch.original_node_ids()].contiguous()
        grads = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes]
        masks, keep = _mask_struct(model)
        gs = _lib.GnnsegGrads(*[g.data_ptr() for g in grads[:10]])
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_backward_nodes(_ptr(ctx.blob), _ptr(ctx.head), C.byref(masks), C.byref(batch.struct), batch.F,
                                               model.hidden_dim, model.n_iters, _ptr(dnode), C.byref(gs), _ptr(grads[10]),
                                               _ptr(grads[11]), _ptr(ctx.ws), ctx.ws.numel(), _stream_ptr(dev)),
                       "gnnseg_backward_nodes")
        del keep
        out = [g if need else None for g, need in zip(grads, ctx.needs_input_grad[2:])]
        return (None, None, *out)


__all__ = ["NodeClassifier", "NodeClfFunction", "DeviceGraphBatch"]
