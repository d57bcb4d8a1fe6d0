"""Time GPU segment construction on ACTS-like (4k hits) and mu200-like (100k hits) events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_fpga_b200.segments import build_segments_device
R = np.array([32., 72., 116., 172., 260., 360., 500., 660., 820., 1020.])
for n_tracks, c_in in ((400, 0.0006), (10000, 0.00006)):
    rng = np.random.RandomState(1)
    layer = np.repeat(np.arange(10), n_tracks); track = np.tile(np.arange(n_tracks), 10)
    r = (R[layer] + rng.normal(0, 0.5, layer.shape[0])).astype(np.float32)
    phi = ((rng.uniform(-np.pi, np.pi, n_tracks)[track] + rng.normal(0, 2.5e-4, n_tracks)[track] * r + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    z = (rng.normal(0, 50, n_tracks)[track] + rng.uniform(-1, 1, n_tracks)[track] * r).astype(np.float32)
    dev = torch.device("cuda:0")
    cols = [torch.as_tensor(a).to(dev) for a in (layer.astype(np.int32), r, phi, z, track.astype(np.int64))]
    pairs = np.stack([np.arange(9), np.arange(1, 10)], axis=1)
    for _ in range(2):
        src, dst, y = build_segments_device(cols[0], cols[1], cols[2], cols[3], cols[4], pairs, c_in, 2 * c_in, 200.0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        src, dst, y = build_segments_device(cols[0], cols[1], cols[2], cols[3], cols[4], pairs, c_in, 2 * c_in, 200.0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("hits %6d pair tests %.3g edges %7d true %6d : %.3f ms per event (count + scan + fill, one host read)" %
          (layer.shape[0], 9.0 * n_tracks * n_tracks, src.numel(), int(y.sum().item()), dt * 1e3))

# a 64-event hit table -> one padded device batch (construct_graphs_device): one H2D per column, one host read
from gnn_fpga_b200.segments import construct_graphs_device, construct_graph_device
n_tracks, c_in, B = 400, 0.0006, 64
rng = np.random.RandomState(2)
cols = {k: [] for k in ("evtid", "layer", "r", "phi", "z", "particle_id")}
for e in range(B):
    layer = np.repeat(np.arange(10), n_tracks); track = np.tile(np.arange(n_tracks), 10)
    r = (R[layer] + rng.normal(0, 0.5, layer.shape[0])).astype(np.float32)
    cols["evtid"].append(np.full(layer.shape[0], e, np.int64)); cols["layer"].append(layer.astype(np.int64)); cols["r"].append(r)
    cols["phi"].append(((rng.uniform(-np.pi, np.pi, n_tracks)[track] + rng.normal(0, 2.5e-4, n_tracks)[track] * r + np.pi) % (2 * np.pi) - np.pi).astype(np.float32))
    cols["z"].append((rng.normal(0, 50, n_tracks)[track] + rng.uniform(-1, 1, n_tracks)[track] * r).astype(np.float32))
    cols["particle_id"].append((track + 1000 * e).astype(np.int64))
hits = {k: np.concatenate(v) for k, v in cols.items()}
pairs = np.stack([np.arange(9), np.arange(1, 10)], axis=1)
for _ in range(2):
    batch, y, n_edges = construct_graphs_device(hits, pairs, c_in, c_in, 2 * c_in, 200.0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    batch, y, n_edges = construct_graphs_device(hits, pairs, c_in, c_in, 2 * c_in, 200.0)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("batch of %d events, %d hits, %d edges (e_max %d): %.2f ms per batch = %.3f ms per event (hit table on the host -> DeviceGraphBatch with CSR)"
      % (B, hits["layer"].shape[0], int(n_edges.sum()), batch.e_max, dt * 1e3, dt * 1e3 / B))
