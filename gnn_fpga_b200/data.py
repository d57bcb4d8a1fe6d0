"""Seeded synthetic events shaped like the reference's datasets.

Every generator returns host `SparseGraph` tuples whose index arrays are in np.nonzero
order, i.e. exactly what `make_sparse_graph` (gnn/graph.py:23-26) would give for the dense
matrices, without building the dense matrices.

* toy2d_*   : the straight-track toy of gnn/MPNN_Seg_Toy2D.ipynb (cells 4, 7, 11): 10 layers,
              4 tracks, all pairs on adjacent layers => N=40, E=144.
* acts_like : 10 barrel layers, n_tracks tracks/event, edges between adjacent layers selected
              with the reference's own cut form abs(dphi/dr) < c1 and abs(z0) < c2
              (gnn/graph.py:58-66).  400 tracks => N=4000, E ~ 20k; 10000 tracks => N=100k, E ~ 1M
              ("mu200-like").
"""
import numpy as np

from .graph import SparseGraph

RAND_PER_K = 3.0   # calibrated: random in-window pairs per hit per unit window (z0_max=200)
BARREL_R = np.array([32., 72., 116., 172., 260., 360., 500., 660., 820., 1020.])  # mm


def sparse_from_edges(X, src, dst, y):
    """Endpoints per edge (edge id = position) -> SparseGraph in np.nonzero order."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    eid = np.arange(src.shape[0], dtype=np.int64)
    oi = np.lexsort((eid, dst))
    oo = np.lexsort((eid, src))
    return SparseGraph(np.ascontiguousarray(X, dtype=np.float32), dst[oi], eid[oi], src[oo], eid[oo],
                       np.asarray(y, dtype=np.float32))


def toy2d_graphs(n_events=32, n_tracks=4, n_layers=10, input_dim=3, seed=0):
    rng = np.random.RandomState(seed)
    det_r = np.arange(n_layers, dtype=np.float64)
    layer = np.repeat(np.arange(n_layers), n_tracks)              # node -> layer
    d = layer[None, :] - layer[:, None]
    src, dst = np.where(d == 1)                                   # all pairs on adjacent layers
    graphs = []
    for _ in range(n_events):
        xin = rng.uniform(size=n_tracks).astype(np.float32)
        xout = rng.uniform(size=n_tracks).astype(np.float32)
        slopes = (xout - xin) / (det_r[-1] - det_r[0])
        x = np.outer(slopes, det_r) + xin[:, None]                # (track, layer)
        order = np.argsort(x, axis=0)                             # hits sorted per layer
        xs = np.take_along_axis(x, order, axis=0).T.reshape(-1)   # (layer, hit) flattened
        label = order.T.reshape(-1)
        feats = [xs, det_r[layer] / n_layers, layer / float(n_layers)][:input_dim]
        X = np.stack(feats, axis=-1).astype(np.float32)
        y = (label[src] == label[dst]).astype(np.float32)
        graphs.append(sparse_from_edges(X, src, dst, y))
    return graphs


def _window_pairs(phi1, phi2, half_width):
    """All (i, j) with |wrap(phi2[j] - phi1[i])| < half_width, via a sorted window join.
    Returned i-major, j ascending (the order of the reference's pandas merge + filter)."""
    order2 = np.argsort(phi2, kind="stable")
    p2 = phi2[order2]
    ext = np.concatenate([p2 - 2 * np.pi, p2, p2 + 2 * np.pi])
    ext_idx = np.concatenate([order2, order2, order2])
    lo = np.searchsorted(ext, phi1 - half_width, side="right")
    hi = np.searchsorted(ext, phi1 + half_width, side="left")
    cnt = np.maximum(hi - lo, 0)
    i = np.repeat(np.arange(phi1.shape[0]), cnt)
    offs = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    j = ext_idx[np.repeat(lo, cnt) + offs]
    key = np.lexsort((j, i))
    return i[key], j[key]


def acts_like_graph(n_tracks=400, seed=0, edges_per_hit=5.0, z0_max=200.0):
    """One event with ~edges_per_hit * 10 * n_tracks edges: the phi-slope window is sized
    from the hit density so that true segments (0.9 per hit) plus random in-window pairs
    reach the target.  400 tracks, 5.0 => N=4000, E ~ 20k (BASELINE.json configs[1])."""
    rng = np.random.RandomState(seed)
    L = BARREL_R.shape[0]
    # random pairs per hit = RAND_PER_K * k for a window phi_slope_max = k * 16 pi / n / 100
    phi_slope_max = max(edges_per_hit - 0.9, 0.1) / RAND_PER_K * 16.0 * np.pi / n_tracks / 100.0
    phi0 = rng.uniform(-np.pi, np.pi, n_tracks)
    kappa = rng.normal(0.0, min(2.5e-4, 0.4 * phi_slope_max), n_tracks)
    z0 = rng.normal(0.0, 50.0, n_tracks)
    cot = rng.uniform(-1.0, 1.0, n_tracks)
    r = np.repeat(BARREL_R, n_tracks)                                     # layer-major nodes
    track = np.tile(np.arange(n_tracks), L)
    phi = phi0[track] + kappa[track] * r
    phi = (phi + np.pi) % (2 * np.pi) - np.pi
    z = z0[track] + cot[track] * r
    X = np.stack([r / 1000.0, phi / np.pi, z / 1000.0], axis=-1).astype(np.float32)
    srcs, dsts = [], []
    for l in range(L - 1):
        a = slice(l * n_tracks, (l + 1) * n_tracks)
        b = slice((l + 1) * n_tracks, (l + 2) * n_tracks)
        dr = BARREL_R[l + 1] - BARREL_R[l]
        i, j = _window_pairs(phi[a], phi[b], phi_slope_max * dr)
        dz = z[b][j] - z[a][i]
        zz0 = z[a][i] - BARREL_R[l] * dz / dr
        keep = np.abs(zz0) < z0_max
        srcs.append(i[keep] + l * n_tracks)
        dsts.append(j[keep] + (l + 1) * n_tracks)
    src = np.concatenate(srcs)
    dst = np.concatenate(dsts)
    y = (track[src] == track[dst]).astype(np.float32)
    return sparse_from_edges(X, src, dst, y)


def acts_like_graphs(n_events=64, n_tracks=400, seed=0, **kw):
    return [acts_like_graph(n_tracks, seed + b, **kw) for b in range(n_events)]


def mu200_like_graph(seed=0, n_tracks=10000, edges_per_hit=10.0):
    """~100k hits, ~1M edges (BASELINE.json configs[3])."""
    return acts_like_graph(n_tracks=n_tracks, seed=seed, edges_per_hit=edges_per_hit)


def random_masks(input_dim, hidden_dim, keep=0.5, seed=1234):
    """Bernoulli(keep) 0/1 masks shaped like masks_e / masks_n
    (gnn/MPNN_Seg_ACTS_maskedlinear.ipynb cell 21)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    D = input_dim + hidden_dim
    shapes = [(hidden_dim, 2 * D), (1, hidden_dim), (hidden_dim, 3 * D), (hidden_dim, hidden_dim)]
    m = [(torch.rand(s, generator=g) < keep).float() for s in shapes]
    return m[:2], m[2:]
