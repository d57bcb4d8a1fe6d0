"""Segment (edge candidate) construction on the GPU: the drop-in for construct_graph /
construct_segments / select_segments of gnn/graph.py:44-142, the step in front of the classifier.

The reference builds, per layer pair, the pandas cross join of the hits of both layers and filters
it with the phi-slope and z0 cuts (7.7 s for two mu200 events, GraphConstructionDev_mu200.ipynb cell
18).  Here the same cuts run in a two-pass CUDA kernel (count, scan, fill: gnnseg_build_segments)
in the columns' own dtype and the reference's order of evaluation, so the kept pairs, their order
and the labels are bit-identical.  There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .graph import DeviceGraphBatch, SparseGraph, _ptr, _require_cuda, _stream_ptr


def build_segments_device(layer, r, phi, z, particle_id, layer_pairs, phi_slope_max, phi_slope_outer_max, z0_max,
                          outer_from_layer=5, device="cuda"):
    """Hit columns (numpy arrays or tensors; r/phi/z all float32 or all float64) -> (src, dst, y)
    int32 / int32 / float32 CUDA tensors of the kept (start, end) positional hit indices in the
    reference's edge order.  One host read of the edge count between the two passes."""
    L = _lib.lib()
    dev = _require_cuda(torch.device(device))
    t = lambda a, dt=None: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(device=dev, dtype=dt).contiguous()
    r, phi, z = t(r), t(phi), t(z)
    if not (r.dtype == phi.dtype == z.dtype and r.dtype in (torch.float32, torch.float64)):
        raise ValueError("r, phi, z must share one dtype, float32 or float64 (got %s, %s, %s)" % (r.dtype, phi.dtype, z.dtype))
    layer = t(layer, torch.int32)
    pid = t(particle_id, torch.int64) if particle_id is not None else None
    n = int(r.numel())
    pairs = np.ascontiguousarray(np.asarray(layer_pairs, dtype=np.int32).reshape(-1, 2))
    n_layers = int(max(int(pairs.max()) + 1 if pairs.size else 1, int(layer.max().item()) + 1 if n else 1))
    if n_layers > 32 or pairs.shape[0] > 32:
        raise ValueError("at most 32 layers and 32 layer pairs")
    wsb = L.gnnseg_segments_workspace_bytes(n, pairs.shape[0])
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    n_edges = torch.zeros(2, dtype=torch.int32, device=dev)
    nb = 4 if r.dtype == torch.float32 else 8

    def call(cap, src, dst, y):
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_build_segments(_ptr(layer), _ptr(r), _ptr(phi), _ptr(z), nb, _ptr(pid) if pid is not None else None,
                                               n, pairs.ctypes.data, pairs.shape[0], n_layers, float(phi_slope_max),
                                               float(phi_slope_outer_max), float(z0_max), int(outer_from_layer), 0, cap,
                                               _ptr(src) if src is not None else None, _ptr(dst) if dst is not None else None,
                                               _ptr(y) if y is not None else None, _ptr(n_edges), _ptr(ws), wsb,
                                               _stream_ptr(dev)), "gnnseg_build_segments")

    call(0, None, None, None)                                   # count
    m = int(n_edges[0].item())
    src = torch.empty(m, dtype=torch.int32, device=dev)
    dst = torch.empty(m, dtype=torch.int32, device=dev)
    y = torch.empty(m, dtype=torch.float32, device=dev)
    if m:
        call(m, src, dst, y)                                    # fill
    return src, dst, y


def scale_features_device(cols, feature_scale, device="cuda"):
    """X = (hits[feature_names].values / feature_scale).astype(np.float32) (gnn/graph.py:118) for three columns."""
    L = _lib.lib()
    dev = _require_cuda(torch.device(device))
    cols = [torch.as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c).to(dev).contiguous() for c in cols]
    if len(cols) != 3 or not all(c.dtype == cols[0].dtype for c in cols) or cols[0].dtype not in (torch.float32, torch.float64):
        raise ValueError("three feature columns of one dtype (float32 or float64) are supported")
    n = int(cols[0].numel())
    X = torch.empty((n, 3), dtype=torch.float32, device=dev)
    s = [float(v) for v in np.asarray(feature_scale, dtype=np.float64)]
    with torch.cuda.device(dev):
        _lib.check(L.gnnseg_scale_features(_ptr(cols[0]), _ptr(cols[1]), _ptr(cols[2]), 4 if cols[0].dtype == torch.float32 else 8,
                                           n, s[0], s[1], s[2], _ptr(X), _stream_ptr(dev)), "gnnseg_scale_features")
    return X


def construct_graph_device(hits, layer_pairs, phi_slope_max, phi_slope_mid_max, phi_slope_outer_max, z0_max,
                           feature_names=("r", "phi", "z"), feature_scale=(1000., np.pi / 8, 1000.), device="cuda"):
    """construct_graph (gnn/graph.py:100-142) for one event, on the device.  `hits`: a pandas
    DataFrame or a dict of columns with layer, r, phi, z, particle_id and the feature columns.
    `phi_slope_mid_max` is accepted and unused, as in the reference (gnn/graph.py:65).
    Returns (batch, y): a one-event DeviceGraphBatch ready for SegmentClassifier(...) and the
    float32 labels per edge; batch.to_sparse_graph(y) gives the reference's SparseGraph tuple."""
    col = lambda k: (hits[k].values if hasattr(hits[k], "values") else hits[k])
    src, dst, y = build_segments_device(col("layer"), col("r"), col("phi"), col("z"),
                                        col("particle_id") if "particle_id" in hits else None, layer_pairs,
                                        phi_slope_max, phi_slope_outer_max, z0_max, device=device)
    X = scale_features_device([col(k) for k in feature_names], feature_scale, device=device)
    batch = DeviceGraphBatch(X, src, dst, 1, int(src.numel()), n_nodes_per_event=[int(X.shape[0])])
    return batch, y


def sparse_graph_of(batch, y=None):
    """The reference's SparseGraph tuple (np.nonzero order, gnn/graph.py:23-26) of a one-event device batch."""
    n_in, n_out = int(batch.in_ptr[-1]), int(batch.out_ptr[-1])
    rows = lambda ptr: np.repeat(np.arange(batch.n_nodes, dtype=np.int64), np.diff(ptr.cpu().numpy()))
    return SparseGraph(batch.X.cpu().numpy(), rows(batch.in_ptr), batch.in_eid[:n_in].cpu().numpy().astype(np.int64),
                       rows(batch.out_ptr), batch.out_eid[:n_out].cpu().numpy().astype(np.int64),
                       y.cpu().numpy() if y is not None else np.zeros(batch.n_slots, np.float32))
