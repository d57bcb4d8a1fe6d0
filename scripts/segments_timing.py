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
