"""GPU segment construction (gnnseg_build_segments / gnnseg_scale_features through the C ABI)
against the reference's own construct_graph outputs (tests/golden/segments_*.npz) and the numpy
oracle: kept pairs, their order, labels, features and the resulting SparseGraph bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import segments_oracle as S

pytestmark = pytest.mark.gpu
CASES = ["segments_f32_small", "segments_f32_event", "segments_f64_small", "segments_f32_wide"]


@pytest.mark.parametrize("name", CASES)
def test_segments_match_reference_construct_graph(name, cuda_device):
    from gnn_fpga_b200.segments import construct_graph_device, sparse_graph_of
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    hits = {k: z[k] for k in ("layer", "r", "phi", "z", "particle_id")}
    c = float(z["phi_slope_max"])
    batch, y = construct_graph_device(hits, z["layer_pairs"], c, c, float(z["phi_slope_outer_max"]), float(z["z0_max"]),
                                      feature_scale=z["feature_scale"], device=cuda_device)
    assert batch.n_slots == z["seg_start"].shape[0]
    assert np.array_equal(batch.src.cpu().numpy(), z["seg_start"]) and np.array_equal(batch.dst.cpu().numpy(), z["seg_end"])
    assert np.array_equal(y.cpu().numpy(), z["y"])
    assert np.array_equal(batch.X.cpu().numpy(), z["X"])
    sg = sparse_graph_of(batch, y)                       # make_sparse_graph's tuple (gnn/graph.py:23-26, 140)
    for k in ("Ri_rows", "Ri_cols", "Ro_rows", "Ro_cols"):
        assert np.array_equal(getattr(sg, k), z[k]), k


def test_larger_event_against_oracle_and_into_the_classifier(cuda_device):
    """~2000 tracks (20k hits): the cross join the oracle still finishes in seconds; the device batch
    goes straight into the classifier and scores like the host-built graph."""
    from gnn_fpga_b200 import SegmentClassifier, SparseGraph
    from gnn_fpga_b200.segments import construct_graph_device, sparse_graph_of
    rng = np.random.RandomState(7)
    R = np.array([32., 72., 116., 172., 260., 360., 500., 660., 820., 1020.])
    n_tracks = 2000
    layer = np.repeat(np.arange(10), n_tracks)
    track = np.tile(np.arange(n_tracks), 10)
    perm = rng.permutation(layer.shape[0])
    layer, track = layer[perm], track[perm]
    r = (R[layer] + rng.normal(0, 0.5, layer.shape[0])).astype(np.float32)
    phi0, kappa = rng.uniform(-np.pi, np.pi, n_tracks), rng.normal(0, 2.5e-4, n_tracks)
    phi = ((phi0[track] + kappa[track] * r + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    zz = (rng.normal(0, 50, n_tracks)[track] + rng.uniform(-1, 1, n_tracks)[track] * r).astype(np.float32)
    pairs = np.stack([np.arange(9), np.arange(1, 10)], axis=1)
    hits = {"layer": layer.astype(np.int32), "r": r, "phi": phi, "z": zz, "particle_id": track.astype(np.int64)}
    batch, y = construct_graph_device(hits, pairs, 0.0004, 0.0004, 0.0008, 200.0, device=cuda_device)
    s, e, yr = S.build_segments(layer, r, phi, zz, track, pairs, 0.0004, 0.0008, 200.0)
    assert np.array_equal(batch.src.cpu().numpy(), s) and np.array_equal(batch.dst.cpu().numpy(), e)
    assert np.array_equal(y.cpu().numpy(), yr) and s.shape[0] > 20000
    torch.manual_seed(0)
    model = SegmentClassifier(3, 32, 2).to(cuda_device).eval()
    with torch.no_grad():
        a = model(batch).clone()
        b = model([sparse_graph_of(batch, y)])
    assert torch.equal(a, b)


def test_count_only_overflow_and_errors(cuda_device):
    import ctypes as C
    from gnn_fpga_b200 import _lib
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    L = _lib.lib()
    z = np.load(os.path.join(GOLDEN, "segments_f32_small.npz"))
    dev = cuda_device
    t = lambda a, dt: torch.as_tensor(a).to(device=dev, dtype=dt).contiguous()
    layer, r, phi, zz = t(z["layer"], torch.int32), t(z["r"], torch.float32), t(z["phi"], torch.float32), t(z["z"], torch.float32)
    n, pairs = r.numel(), np.ascontiguousarray(z["layer_pairs"].astype(np.int32))
    wsb = L.gnnseg_segments_workspace_bytes(n, pairs.shape[0])
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ne = torch.zeros(2, dtype=torch.int32, device=dev)
    m = z["seg_start"].shape[0]
    args = lambda cap, s, d: (_ptr(layer), _ptr(r), _ptr(phi), _ptr(zz), 4, None, n, pairs.ctypes.data, pairs.shape[0], 10,
                              float(z["phi_slope_max"]), float(z["phi_slope_outer_max"]), float(z["z0_max"]), 5, 100, cap,
                              _ptr(s) if s is not None else None, _ptr(d) if d is not None else None, None, _ptr(ne), _ptr(ws), wsb,
                              _stream_ptr(dev))
    assert L.gnnseg_build_segments(*args(0, None, None)) == 0
    assert ne.tolist() == [m, 1]                                            # counted; capacity 0 is "too small"
    src = torch.full((m,), -7, dtype=torch.int32, device=dev)
    dst = torch.full((m,), -7, dtype=torch.int32, device=dev)
    assert L.gnnseg_build_segments(*args(m // 2, src, dst)) == 0            # half the room: prefix written, flag set
    assert ne.tolist() == [m, 1]
    assert np.array_equal(src[:m // 2].cpu().numpy(), z["seg_start"][:m // 2] + 100) and torch.all(src[m // 2:] == -7)
    assert L.gnnseg_build_segments(*args(m, src, dst)) == 0
    assert ne.tolist() == [m, 0] and np.array_equal(dst.cpu().numpy(), z["seg_end"] + 100)    # node_offset = 100
    bad = list(args(m, src, dst)); bad[4] = 2
    assert L.gnnseg_build_segments(*bad) == -2                              # dtype
    bad = list(args(m, src, dst)); bad[9] = 5
    assert L.gnnseg_build_segments(*bad) == -1                              # a pair names layer >= n_layers
    bad = list(args(m, src, dst)); bad[21] = 64
    assert L.gnnseg_build_segments(*bad) == -3                              # workspace


@pytest.mark.parametrize("name", ["segments_multi_plain", "segments_multi_tracks", "segments_multi_nomiss"])
def test_event_batches_match_reference_event_loop(name, cuda_device):
    """construct_graphs_device on a multi-event hit table (rows of the events interleaved) against the
    reference's event loop around construct_graph: one padded device batch, bit-exact per event."""
    from gnn_fpga_b200 import SegmentClassifier, SparseGraph
    from gnn_fpga_b200.segments import construct_graphs_device
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    hits = {k: z["hits_" + k] for k in ("evtid", "layer", "r", "phi", "z", "particle_id")}
    mt = int(z["max_tracks"])
    np.random.seed(int(z["seed"]))
    c = float(z["phi_slope_max"])
    batch, y, n_edges = construct_graphs_device(hits, z["layer_pairs"], c, c, float(z["phi_slope_outer_max"]), float(z["z0_max"]),
                                                feature_scale=z["feature_scale"], max_tracks=None if mt < 0 else mt,
                                                no_missing_hits=bool(int(z["no_missing_hits"])), device=cuda_device)
    B = len(z["n_hits"])
    assert batch.B == B and list(batch.n_nodes_per_event) == list(z["n_hits"]) and list(n_edges) == list(z["n_edges"])
    e_max = int(max(z["n_edges"]))
    assert batch.e_max == e_max and y.shape == (B, e_max)
    src, dst = batch.src.cpu().numpy().reshape(B, e_max), batch.dst.cpu().numpy().reshape(B, e_max)
    off = np.concatenate([[0], np.cumsum(z["n_hits"])])
    graphs = []
    for b in range(B):
        m = int(z["n_edges"][b])
        assert np.array_equal(src[b, :m], z["src_%d" % b] + off[b]) and np.array_equal(dst[b, :m], z["dst_%d" % b] + off[b])
        assert np.all(src[b, m:] == -1) and np.all(dst[b, m:] == -1)
        assert np.array_equal(y[b, :m].cpu().numpy(), z["y_%d" % b]) and float(y[b, m:].abs().sum()) == 0
        assert np.array_equal(batch.X[off[b]:off[b + 1]].cpu().numpy(), z["X_%d" % b])
        eid = np.arange(m, dtype=np.int64)
        oi, oo = np.lexsort((eid, z["dst_%d" % b])), np.lexsort((eid, z["src_%d" % b]))
        graphs.append(SparseGraph(z["X_%d" % b], z["dst_%d" % b][oi], eid[oi], z["src_%d" % b][oo], eid[oo], z["y_%d" % b]))
    # the device-built batch scores exactly like the same events packed from host tuples
    torch.manual_seed(0)
    model = SegmentClassifier(3, 32, 2).to(cuda_device).eval()
    with torch.no_grad():
        a = model(batch).clone()
        b_ = model(graphs)
    assert torch.equal(a, b_)
