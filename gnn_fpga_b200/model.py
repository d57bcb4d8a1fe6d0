"""Drop-in SegmentClassifier whose forward runs on the B200 CUDA path.

Module tree, constructor signature, parameter names/shapes/order and `state_dict` keys are
those of the reference (gnn/model.py:14-156), so reference checkpoints load and
`Estimator` (gnn/estimator.py) can drive it unchanged.  What differs is the execution:
`SegmentClassifier.forward` packs the (masked) weights and calls `gnnseg_forward` through
the C ABI (include/gnnseg.h); no dense incidence bmm ever runs.

The sub-modules keep their reference attributes (`network[0].weight`, `.mask`,
`.mask_flag`, `set_mask`, `get_mask`) because estimator.py:54-55 and
estimator_maskedlinear.py:85-101 reach into them.  Their own `forward` methods are not
implemented: the fused path never calls them and there is no CPU fallback.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from .graph import DeviceGraphBatch, SparseGraph, _ptr, _stream_ptr


class MaskedLinear(nn.Linear):
    """nn.Linear with an optional 0/1 weight mask (gnn/model.py:14-33).  As in the
    reference the mask is a plain attribute: not a buffer, not in state_dict."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias)
        self.mask_flag = False
        self.mask = None

    def set_mask(self, mask):
        # gnn/model.py:19-22 (and the twin's device handling, model_maskedlinear.py:19-30):
        # the mask follows the weight's device, the stored weights are zeroed where masked.
        self.mask = mask.to(self.weight.device) if isinstance(mask, torch.Tensor) else torch.as_tensor(mask, device=self.weight.device)
        self.weight.data = self.weight.data * self.mask.data
        self.mask_flag = True

    def get_mask(self):
        return self.mask

    def effective_mask(self):
        """Mask tensor on the weight's device (fp32, contiguous) or None."""
        if not self.mask_flag or self.mask is None:
            return None
        m = self.mask.detach()
        if m.device != self.weight.device or m.dtype != torch.float32 or not m.is_contiguous():
            m = m.to(device=self.weight.device, dtype=torch.float32).contiguous()
            self.mask = m
        return m

    def forward(self, x):
        raise _lib.GnnsegError("MaskedLinear.forward is not a standalone op here: call SegmentClassifier(...)")


def _is_file_batch(inputs):
    """A batch given as graph file names (the reference's NPZ format) instead of tuples."""
    return isinstance(inputs, (list, tuple)) and len(inputs) > 0 and isinstance(inputs[0], (str, os.PathLike))


def _check_activation(act):
    if act is not nn.Tanh:
        raise ValueError("hidden_activation=%r: the CUDA path implements nn.Tanh only "
                         "(every reference configuration uses Tanh)" % (act,))


class EdgeNetwork(nn.Module):
    """Parameter container matching gnn/model.py:36-55."""

    def __init__(self, input_dim, hidden_dim=8, hidden_activation=nn.Tanh, mask=None):
        super().__init__()
        _check_activation(hidden_activation)
        self.network = nn.Sequential(
            MaskedLinear(input_dim * 2, hidden_dim), hidden_activation(),
            MaskedLinear(hidden_dim, 1), nn.Sigmoid())
        self.mask = mask
        if mask is not None:
            self.network[0].set_mask(mask[0])
            self.network[2].set_mask(mask[1])

    def forward(self, *a, **k):
        raise _lib.GnnsegError("EdgeNetwork runs fused inside SegmentClassifier.forward (gnnseg_edge_step)")


class NodeNetwork(nn.Module):
    """Parameter container matching gnn/model.py:84-101.  Unlike gnn/model.py:100 (which
    raises TypeError) `mask=None` is accepted, as in model_maskedlinear.py:110-112."""

    def __init__(self, input_dim, output_dim, hidden_activation=nn.Tanh, mask=None):
        super().__init__()
        _check_activation(hidden_activation)
        self.network = nn.Sequential(
            MaskedLinear(input_dim * 3, output_dim), hidden_activation(),
            MaskedLinear(output_dim, output_dim), hidden_activation())
        self.mask = mask
        if mask is not None:
            self.network[0].set_mask(mask[0])
            self.network[2].set_mask(mask[1])

    def forward(self, *a, **k):
        raise _lib.GnnsegError("NodeNetwork runs fused inside SegmentClassifier.forward (gnnseg_node_step)")


class SegmentClassifier(nn.Module):
    """gnn/model.py:127-156.  `forward(inputs)` accepts

    * `[X, Ri, Ro]`: dense fp32 CUDA tensors (B,N,F), (B,N,E), (B,N,E) -> (B,E) scores, the
      reference call;
    * a `DeviceGraphBatch` -> (B, e_max) scores (replayed from a CUDA graph on repeat calls);
    * a list of host `SparseGraph` tuples -> (B, e_max) scores, padded as merge_graphs pads;
    * a list of graph file names (the NPZ format of gnn/graph.py:179-194) -> the same, the files
      being mapped, parsed and packed by the library (no np.load).
    """

    def __init__(self, input_dim=2, hidden_dim=8, n_iters=3, hidden_activation=nn.Tanh,
                 masks_e=None, masks_n=None):
        super().__init__()
        _check_activation(hidden_activation)
        self.n_iters = n_iters
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.input_network = nn.Sequential(nn.Linear(input_dim, hidden_dim), hidden_activation())
        self.edge_network = EdgeNetwork(input_dim + hidden_dim, hidden_dim, hidden_activation, masks_e)
        self.node_network = NodeNetwork(input_dim + hidden_dim, hidden_dim, hidden_activation, masks_n)
        self._blob = None
        self.exact = False           # True: always the step-by-step kernels (GNNSEG_FWD_EXACT), see check_range
        self._range_checks = []      # (event, pinned status word) of forwards whose range flag has not been looked at
        self._stream_cache = None    # predict_stream's pipeline slots and side streams, kept between calls
        self.use_cuda_graph = True
        self._arena = None           # reusable pinned arena for batches given as host SparseGraph tuples

    # -- weights -------------------------------------------------------------------------
    def _device(self):
        return self.input_network[0].weight.device

    def _params_struct(self):
        """GnnsegParams over the ten parameter tensors and four masks (+ the tensors to keep alive)."""
        e0, e2 = self.edge_network.network[0], self.edge_network.network[2]
        n0, n2 = self.node_network.network[0], self.node_network.network[2]
        tensors = [self.input_network[0].weight, self.input_network[0].bias,
                   e0.weight, e0.bias, e2.weight, e2.bias, n0.weight, n0.bias, n2.weight, n2.bias]
        keep = []
        ptrs = []
        for t in tensors:
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(torch.float32).contiguous()
            keep.append(t)
            ptrs.append(t.data_ptr())
        for lin in (e0, e2, n0, n2):
            m = lin.effective_mask()
            keep.append(m)
            ptrs.append(m.data_ptr() if m is not None else None)
        return _lib.GnnsegParams(*ptrs), keep

    def pack_weights(self):
        """W*mask, transposed and padded into the kernel layout (gnnseg_pack_weights).  Runs
        every forward, like the reference recomputes weight*mask (gnn/model.py:30)."""
        L = _lib.lib()
        dev = self._device()
        if dev.type != "cuda":
            raise _lib.GnnsegError("SegmentClassifier parameters are on %s: move the model to a CUDA "
                                   "device (model.cuda()); there is no CPU path" % dev)
        F, h = self.input_dim, self.hidden_dim
        n = L.gnnseg_weights_floats(F, h)
        if n == 0:
            _lib.check(-2, "SegmentClassifier(input_dim=%d, hidden_dim=%d)" % (F, h))
        if self._blob is None or self._blob.device != dev or self._blob.numel() != n:
            self._blob = torch.empty(n, dtype=torch.float32, device=dev)
        params, keep = self._params_struct()
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_pack_weights(C.byref(params), F, h, _ptr(self._blob), _stream_ptr(dev)),
                       "gnnseg_pack_weights")
        del keep
        return self._blob

    # -- forward -------------------------------------------------------------------------
    def _run(self, batch):
        L = _lib.lib()
        dev = batch.device
        if batch.F != self.input_dim:
            raise ValueError("X has %d features, model expects input_dim=%d" % (batch.F, self.input_dim))
        blob = self.pack_weights()
        if blob.device != dev:
            raise ValueError("graph batch on %s but model on %s" % (dev, blob.device))
        h = self.hidden_dim
        ws = batch.workspace(h)

        flags = _lib.FWD_EXACT if self.exact else 0

        def launch():
            _lib.check(L.gnnseg_forward_ex(_ptr(blob), C.byref(batch.struct), _ptr(batch.X), batch.F, h,
                                           self.n_iters, _ptr(batch.scores), _ptr(ws), ws.numel(), flags,
                                           _ptr(batch.status), _stream_ptr(dev)), "gnnseg_forward_ex")

        with torch.cuda.device(dev):
            if not self.use_cuda_graph:
                launch()
                return batch.scores
            key = (blob.data_ptr(), h, self.n_iters)
            entry = batch._graphs.get(key)
            if entry is None:
                # first call with this batch: plain launches (also warms up the kernels)
                launch()
                batch._graphs[key] = "warm"
            elif entry == "warm":
                # second call: capture the 2*n_iters+2 launches once, then replay
                torch.cuda.current_stream(dev).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    launch()
                batch._graphs[key] = graph
                graph.replay()
            else:
                entry.replay()
        return batch.scores

    def _tuple_store(self, graphs, arena=None, n_threads=0):
        """Host SparseGraph tuples -> a one-batch GraphStore (validated: ValueError for out-of-range indices
        or a column listed twice).  Node order kept: renumbering is load-time work (GraphStore)."""
        from .store import GraphStore
        return GraphStore.from_sparse_graphs(graphs, reorder=False, n_threads=n_threads, arena=arena)

    def predict_stream(self, batches, depth=3):
        """Pipelined inference over an iterable of batches: yields, in order, one pinned host tensor
        (B, E_max) of scores per batch.  A batch is

        * a `StoreBatch` (events [lo, hi) of a GraphStore, gnn_fpga_b200/store.py): the fast path.  The
          host does no per-batch work: five contiguous asynchronous copies out of the store's pinned arena on
          a copy stream, gnnseg_assemble_batch + the forward on the compute stream, the scores back on a
          third stream; `depth` batches are in flight, so the copies of neighbouring batches overlap the
          compute of this one.  Replaces the reference's per-batch graph_from_sparse + merge_graphs + .cuda()
          (gnn/trainSegmentClassifier.py:97-111) and Estimator.predict's loop (gnn/estimator.py:137-146);
        * a list of host SparseGraph tuples or of graph file names: packed on a worker thread into a
          one-batch store first (the per-batch host work the reference does, minus the densifying), then the
          same device path.

        A yielded tensor is reused `depth` batches later."""
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        from .graph import DeviceBatchBuffers, _require_cuda, load_graphs_mapped
        from .store import StoreBatch
        dev = _require_cuda(self._device())
        depth = max(1, int(depth))
        # pipeline slots (device buffers, pinned result buffers, pinned arenas for tuple batches) and the two side
        # streams are kept on the model between calls: allocating pinned memory costs more than a batch
        cache = self._stream_cache if self._stream_cache is not None and self._stream_cache["dev"] == dev else None
        if cache is None:
            cache = self._stream_cache = {"dev": dev, "slots": [], "streams": (torch.cuda.Stream(dev), torch.cuda.Stream(dev)),
                                          "assemble": torch.cuda.Stream(dev)}
        while len(cache["slots"]) < depth + 1:
            cache["slots"].append({"bufs": None, "done": None, "view": None, "flag": None, "arena": None, "h2d": None})
        slots = cache["slots"][:depth + 1]
        for s_ in slots:
            if s_["done"] is not None:
                s_["done"].synchronize()                          # a previous stream that was abandoned half way
        was_graph, self.use_cuda_graph = self.use_cuda_graph, False    # one-shot batches: plain launches
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))     # ranks sharing this host (torchrun)
        pack_threads = max(1, min(4, (os.cpu_count() or 4) // local_world - 2))
        if os.environ.get("GNNSEG_PACK_THREADS"):
            pack_threads = max(1, int(os.environ["GNNSEG_PACK_THREADS"]))

        def prepare(item, slot):
            """Anything that is not a StoreBatch becomes one (worker thread; the C++ fill releases the GIL)."""
            if isinstance(item, StoreBatch):
                return item
            if slot["h2d"] is not None:
                slot["h2d"].synchronize()                                # the slot's last copies have left its arena
            graphs = load_graphs_mapped([os.fspath(f) for f in item]) if _is_file_batch(item) else list(item)
            st = self._tuple_store(graphs, arena=slot["arena"], n_threads=pack_threads)
            slot["arena"] = st.arena
            return StoreBatch(st, 0, len(st))

        def range_ok(slot):
            if int(slot["flag"][0]) & _lib.RANGE_CLAMPED:
                raise _lib.GnnsegError("an edge-network projection exceeded the fused path's range (|W1.[H|X] + b1| > 43.6); "
                                       "set model.exact = True and rerun")

        pool = ThreadPoolExecutor(max_workers=1)
        compute = torch.cuda.current_stream(dev)
        s_in, s_out = cache["streams"]
        L = _lib.lib()
        h = self.hidden_dim
        if L.gnnseg_supported(self.input_dim, h) == 0:
            _lib.check(-2, "SegmentClassifier(input_dim=%d, hidden_dim=%d)" % (self.input_dim, h))
        blob = self.pack_weights()                                # once: the weights do not change while the stream runs
        flags = _lib.FWD_EXACT if self.exact else 0
        import numpy as np
        shape = np.zeros(4, dtype=np.int32)
        pending = deque()
        # GNNSEG_STREAM_TIMING=1: host seconds spent waiting for results / inside the library call, per stream (diagnostics)
        import time
        stats = self._stream_stats = {"wait": 0.0, "call": 0.0, "batches": 0} if os.environ.get("GNNSEG_STREAM_TIMING") else None
        try:
            with torch.no_grad():
                it = iter(batches)
                futs = deque()
                n_sub = 0

                def submit_next():
                    nonlocal n_sub
                    try:
                        item = next(it)
                    except StopIteration:
                        return
                    slot = slots[n_sub % len(slots)]
                    futs.append(item if isinstance(item, StoreBatch) else pool.submit(prepare, item, slot))
                    n_sub += 1

                submit_next()
                i = 0
                while futs:
                    s = slots[i % len(slots)]
                    f = futs.popleft()
                    sb = f if isinstance(f, StoreBatch) else f.result()
                    submit_next()                                 # one batch being prepared ahead
                    if len(pending) == depth:                     # keep `depth` results in flight
                        old = pending.popleft()
                        t0 = time.perf_counter() if stats else 0.0
                        old["done"].synchronize()
                        if stats:
                            stats["wait"] += time.perf_counter() - t0
                        range_ok(old)
                        yield old["view"]
                    if s["done"] is not None:
                        s["done"].synchronize()                   # the slot's previous batch has left the device buffers
                    t0 = time.perf_counter() if stats else 0.0
                    store = sb.store
                    B = sb.hi - sb.lo
                    while True:
                        if s["bufs"] is None:
                            L.gnnseg_store_batch_shape_host(C.byref(store.layout), store.arena.data_ptr(), sb.lo, sb.hi, shape.ctypes.data)
                            n, e_max, n_in, n_out = (int(x) for x in shape)
                            g = lambda v: int(v * 1.25) + 1
                            s["bufs"] = DeviceBatchBuffers(dev, g(n), g(n_in), g(n_out), g(B * e_max), B, store.F,
                                                           store.col_bytes).with_host_results()
                            if os.environ.get("GNNSEG_ASSEMBLE_OWN_STREAM", "1") != "0":     # the batch assembly on a stream of its own
                                s["bufs"].assemble_stream = cache["assemble"].cuda_stream
                        bufs = s["bufs"]
                        rc = _lib.EWORKSPACE
                        if bufs.F == store.F and bufs.col_bytes == store.col_bytes:
                            torch.cuda.set_device(dev)
                            rc = L.gnnseg_store_forward_batch(C.byref(store.layout), store.arena.data_ptr(), sb.lo, sb.hi, _ptr(blob), h,
                                                              self.n_iters, flags, C.byref(bufs.struct(h)), C.c_void_p(s_in.cuda_stream),
                                                              C.c_void_p(compute.cuda_stream), C.c_void_p(s_out.cuda_stream),
                                                              shape.ctypes.data)
                        if rc != _lib.EWORKSPACE:
                            break
                        s["bufs"] = None                          # the batch exceeds the slot's buffers: new ones, once
                    _lib.check(rc, "gnnseg_store_forward_batch")
                    if stats:
                        stats["call"] += time.perf_counter() - t0
                        stats["batches"] += 1
                    e_max = int(shape[1])
                    s["h2d"] = torch.cuda.Event()
                    s["h2d"].record(s_in)
                    s["done"] = torch.cuda.Event()
                    s["done"].record(s_out)
                    s["view"] = bufs.scores_host[:B * e_max].view(B, e_max)
                    s["flag"] = bufs.status_host
                    pending.append(s)
                    i += 1
                while pending:
                    old = pending.popleft()
                    old["done"].synchronize()
                    range_ok(old)
                    yield old["view"]
        finally:
            pool.shutdown(wait=True)
            self.use_cuda_graph = was_graph

    def _watch_range(self, batch):
        """Queue an asynchronous read of the batch's range flag (gnnseg_forward_ex's status word)."""
        host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        host.copy_(batch.status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(batch.device))
        self._range_checks.append((ev, host))

    def check_range(self, wait=True):
        """The fused inference path (hidden_dim 32 / 64) keeps the edge network's first-layer projections as
        exponentials and represents |projection| <= 43.6; a forward that had to clamp one raises its flag.
        Raises GnnsegError if a finished forward (wait=True: any forward issued so far) was flagged; the
        cure is `model.exact = True` (the step-by-step kernels).  predict_stream checks every batch it
        yields; `model(inputs)` checks the calls before it."""
        left = []
        bad = False
        for ev, host in self._range_checks:
            if wait:
                ev.synchronize()
            if ev.query():
                bad |= bool(int(host[0]) & _lib.RANGE_CLAMPED)
            else:
                left.append((ev, host))
        self._range_checks = left
        if bad:
            raise _lib.GnnsegError("an edge-network projection exceeded the fused path's range (|W1.[H|X] + b1| > 43.6): "
                                   "the scores of that forward may be off; set model.exact = True and rerun")

    def _wants_grad(self):
        # training mode + autograd on + trainable parameters: what Estimator.fit_gen / training_step
        # set up (model.train(), gnn/estimator.py:96); eval() or no_grad() always means inference
        return self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _to_batch(self, inputs):
        """Anything forward accepts -> DeviceGraphBatch."""
        from .graph import _require_cuda
        from .store import StoreBatch
        if isinstance(inputs, DeviceGraphBatch):
            return inputs
        if isinstance(inputs, StoreBatch):
            return DeviceGraphBatch.from_store(inputs.store, inputs.lo, inputs.hi, _require_cuda(self._device()))
        if _is_file_batch(inputs):
            return DeviceGraphBatch.from_graph_files(inputs, device=_require_cuda(self._device()))
        if isinstance(inputs, (list, tuple)) and len(inputs) > 0 and isinstance(inputs[0], SparseGraph):
            dev = _require_cuda(self._device())        # fails loudly before any host work: no CPU path
            st = self._tuple_store(list(inputs), arena=self._arena)
            self._arena = st.arena
            batch = DeviceGraphBatch.from_store(st, 0, len(st), dev)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            ev.synchronize()                           # the arena is reused by the next call
            return batch
        X, Ri, Ro = inputs
        if not (isinstance(X, torch.Tensor) and X.is_cuda):
            raise ValueError("SegmentClassifier expects CUDA tensors [X, Ri, Ro] (got %s); there is no CPU path"
                             % (X.device if isinstance(X, torch.Tensor) else type(X)))
        return DeviceGraphBatch.from_dense(X, Ri, Ro)

    def forward(self, inputs):
        """Returns a fresh (B, E_max) tensor per call, like the reference (gnn/model.py:156).  For a
        DeviceGraphBatch the result is a copy of the batch-owned `batch.scores` buffer, which the next call on
        the same batch (or a CUDA-graph replay) overwrites."""
        if self._wants_grad():
            # A forward that must be differentiated (Estimator.training_step, gnn/estimator.py:53):
            # gnnseg_forward_train keeps the activations, the result's grad_fn runs gnnseg_backward
            # (gnn_fpga_b200/training.py).  Inference (eval() or no_grad, as Estimator.predict
            # does) takes the lighter path below.
            from .training import differentiable_forward
            if self._device().type != "cuda":
                raise _lib.GnnsegError("SegmentClassifier parameters are on %s: there is no CPU path" % self._device())
            if _lib.lib().gnnseg_supported(self.input_dim, self.hidden_dim) == 0:   # same shapes as inference
                _lib.check(-2, "SegmentClassifier(input_dim=%d, hidden_dim=%d)" % (self.input_dim, self.hidden_dim))
            return differentiable_forward(self, self._to_batch(inputs))
        resident = isinstance(inputs, DeviceGraphBatch)
        self.check_range(wait=False)
        batch = self._to_batch(inputs)
        out = self._run(batch).view(batch.B, batch.e_max)
        if not self.exact and self.hidden_dim in (32, 64):
            self._watch_range(batch)
        return out.clone() if resident else out        # one-shot batches own their buffer: nothing else writes it
