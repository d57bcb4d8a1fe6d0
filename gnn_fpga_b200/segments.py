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


def _columns(hits, names):
    """Named columns of a pandas DataFrame or of a dict of arrays, as numpy arrays in row order."""
    return {k: np.asarray(hits[k].values if hasattr(hits[k], "values") else hits[k]) for k in names if k in hits}


def select_tracks(particle_id, layer, max_tracks=None, no_missing_hits=False):
    """Row mask of the hit pre-selection at the top of construct_graph (gnn/graph.py:106-114), in
    the same order: first `no_missing_hits` (keep particles seen on exactly 10 distinct layers), then
    `max_tracks` (the distinct particle ids in order of first appearance, shuffled with numpy's
    GLOBAL generator as the reference does, the first max_tracks kept), so that a caller who seeds
    np.random gets the reference's sample."""
    pid = np.asarray(particle_id)
    keep = np.ones(pid.shape[0], dtype=bool)
    if no_missing_hits:
        uniq, inv = np.unique(pid, return_inverse=True)
        seen = np.zeros((uniq.shape[0], int(np.max(layer)) + 1 if pid.size else 1), dtype=bool)
        seen[inv, np.asarray(layer)] = True
        keep &= (seen.sum(axis=1) == 10)[inv]
    if max_tracks is not None:
        sub = pid[keep]
        first = np.sort(np.unique(sub, return_index=True)[1])
        keys = sub[first].copy()
        np.random.shuffle(keys)
        keep &= np.isin(pid, keys[:max_tracks])
    return keep


def event_rows(cols, max_events=None, max_tracks=None, no_missing_hits=False, one_event=False):
    """Host side of the event loop (gnn/graph.py:152-169): row indices of every event of the hit
    table `cols` (dict of numpy columns), events in order of first appearance of their `evtid`, rows
    in table order, after the per-event track pre-selection (select_tracks, drawn event by event
    like the reference's loop).  Pure numpy; no device work."""
    n_all = int(cols["layer"].shape[0])
    if one_event or "evtid" not in cols or n_all == 0:
        groups = [np.arange(n_all)]
    else:
        evt = cols["evtid"]
        cut = np.flatnonzero(evt[1:] != evt[:-1]) + 1                     # starts of the runs of equal ids
        starts = np.concatenate([[0], cut])
        if np.unique(evt[starts]).shape[0] == starts.shape[0]:
            groups = np.split(np.arange(n_all), cut)                      # table already grouped by event: no sort
        else:
            # one stable sort instead of a scan of the table per event: events ranked by the first
            # appearance of their id, rows of an event in table order
            _, first, inv = np.unique(evt, return_index=True, return_inverse=True)
            rank = np.empty(first.shape[0], dtype=np.int64)
            rank[np.argsort(first, kind="stable")] = np.arange(first.shape[0])
            codes = rank[inv.reshape(-1)]
            order = np.argsort(codes, kind="stable")
            groups = np.split(order, np.cumsum(np.bincount(codes, minlength=first.shape[0]))[:-1])
        if max_events is not None:
            groups = groups[:max_events]
    if max_tracks is not None or no_missing_hits:
        if "particle_id" not in cols:
            raise ValueError("max_tracks / no_missing_hits need a particle_id column")
        groups = [g[select_tracks(cols["particle_id"][g], cols["layer"][g], max_tracks, no_missing_hits)] for g in groups]
    return groups


def construct_graph_device(hits, layer_pairs, phi_slope_max, phi_slope_mid_max, phi_slope_outer_max, z0_max,
                           feature_names=("r", "phi", "z"), feature_scale=(1000., np.pi / 8, 1000.), max_tracks=None,
                           no_missing_hits=False, device="cuda"):
    """construct_graph (gnn/graph.py:100-142) for one event, on the device.  `hits`: a pandas
    DataFrame or a dict of columns with layer, r, phi, z, particle_id and the feature columns.
    `phi_slope_mid_max` is accepted and unused, as in the reference (gnn/graph.py:65).
    Returns (batch, y): a one-event DeviceGraphBatch ready for SegmentClassifier(...) and the
    float32 labels per edge; sparse_graph_of(batch, y) gives the reference's SparseGraph tuple."""
    batch, y, _ = construct_graphs_device(hits, layer_pairs, phi_slope_max, phi_slope_mid_max, phi_slope_outer_max, z0_max,
                                          feature_names, feature_scale, max_tracks=max_tracks,
                                          no_missing_hits=no_missing_hits, device=device, _one_event=True)
    return batch, y.view(-1)


def construct_graphs_device(hits, layer_pairs, phi_slope_max, phi_slope_mid_max, phi_slope_outer_max, z0_max,
                            feature_names=("r", "phi", "z"), feature_scale=(1000., np.pi / 8, 1000.), max_events=None,
                            max_tracks=None, no_missing_hits=False, outer_from_layer=5, device="cuda", _one_event=False):
    """The event loop of construct_graphs (gnn/graph.py:145-175: group the hit table by `evtid`, events
    in order of first appearance, construct_graph per event) as ONE padded device batch: the hit
    columns of all selected events go to the device in one copy each, the cuts run per event
    (gnnseg_build_segments_batch: every event in one set of launches, counting pass then filling pass
    straight into the events' slot ranges), one host read (the per-event edge counts) in between.
    Returns (batch, y, n_edges): a DeviceGraphBatch of B events padded to e_max = max edge count
    exactly as merge_graphs pads (absent slots -1), labels (B, e_max) float32 with zeros in the padding,
    and the edge count of every event."""
    L = _lib.lib()
    dev = _require_cuda(torch.device(device))
    names = ["layer", "r", "phi", "z", "particle_id", "evtid"] + list(feature_names)
    cols = _columns(hits, names)
    groups = event_rows(cols, max_events, max_tracks, no_missing_hits, _one_event)
    n_all = int(cols["layer"].shape[0])
    B = len(groups)
    rows = np.concatenate(groups) if B else np.zeros(0, np.int64)
    n_per = [int(g.shape[0]) for g in groups]
    node_off = np.concatenate([[0], np.cumsum(n_per)]).astype(np.int64)
    whole = rows.shape[0] == n_all and np.array_equal(rows, np.arange(n_all))
    take = (lambda a: a) if whole else (lambda a: a[rows])
    to_dev = lambda a, dt=None: torch.as_tensor(np.ascontiguousarray(take(a))).to(device=dev, dtype=dt).contiguous()
    r, phi, z = to_dev(cols["r"]), to_dev(cols["phi"]), to_dev(cols["z"])
    if not (r.dtype == phi.dtype == z.dtype and r.dtype in (torch.float32, torch.float64)):
        raise ValueError("r, phi, z must share one dtype, float32 or float64 (got %s, %s, %s)" % (r.dtype, phi.dtype, z.dtype))
    layer = to_dev(cols["layer"], torch.int32)
    pid = to_dev(cols["particle_id"], torch.int64) if "particle_id" in cols else None
    pairs = np.ascontiguousarray(np.asarray(layer_pairs, dtype=np.int32).reshape(-1, 2))
    n_layers = int(max(int(pairs.max()) + 1 if pairs.size else 1, int(take(cols["layer"]).max()) + 1 if rows.size else 1))
    if n_layers > 32 or pairs.shape[0] > 32:
        raise ValueError("at most 32 layers and 32 layer pairs")
    nb = 4 if r.dtype == torch.float32 else 8
    n_big, n_tot = (max(n_per) if B else 0), int(node_off[-1])
    wsb = L.gnnseg_segments_batch_workspace_bytes(B, n_tot, pairs.shape[0])
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dev)
    counts = torch.zeros((max(B, 1), 2), dtype=torch.int32, device=dev)
    hit_off = torch.as_tensor(node_off.astype(np.int32)).to(dev)

    def call(e_max, src, dst, y):
        opt = lambda t: _ptr(t) if t is not None else None
        _lib.check(L.gnnseg_build_segments_batch(_ptr(layer), _ptr(r), _ptr(phi), _ptr(z), nb, opt(pid), B, _ptr(hit_off), n_tot,
                                                 n_big, pairs.ctypes.data, pairs.shape[0], n_layers, float(phi_slope_max),
                                                 float(phi_slope_outer_max), float(z0_max), int(outer_from_layer), e_max,
                                                 opt(src), opt(dst), opt(y), _ptr(counts), _ptr(ws), wsb, _stream_ptr(dev)),
                   "gnnseg_build_segments_batch")

    with torch.cuda.device(dev):
        call(0, None, None, None)                                   # count: every event in one set of launches
        n_edges = counts[:B, 0].cpu().numpy().astype(np.int64)      # the one host read
        e_max = int(n_edges.max()) if B else 0
        src = torch.full((B * e_max,), -1, dtype=torch.int32, device=dev)
        dst = torch.full((B * e_max,), -1, dtype=torch.int32, device=dev)
        y = torch.zeros((B, e_max), dtype=torch.float32, device=dev)
        if e_max:
            call(e_max, src, dst, y if pid is not None else None)   # fill, on the offsets the count left in ws
    X = scale_features_device([take(cols[k]) for k in feature_names], feature_scale, device=dev)
    batch = DeviceGraphBatch(X, src, dst, B, e_max, n_nodes_per_event=n_per)
    return batch, y, n_edges


def sparse_graph_of(batch, y=None):
    """The reference's SparseGraph tuple (np.nonzero order, gnn/graph.py:23-26) of a one-event device batch."""
    n_in, n_out = int(batch.in_ptr[-1]), int(batch.out_ptr[-1])
    rows = lambda ptr: np.repeat(np.arange(batch.n_nodes, dtype=np.int64), np.diff(ptr.cpu().numpy()))
    return SparseGraph(batch.X.cpu().numpy(), rows(batch.in_ptr), batch.in_eid[:n_in].cpu().numpy().astype(np.int64),
                       rows(batch.out_ptr), batch.out_eid[:n_out].cpu().numpy().astype(np.int64),
                       y.cpu().numpy() if y is not None else np.zeros(batch.n_slots, np.float32))
