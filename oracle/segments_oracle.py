"""CPU oracle for segment construction.  TEST INFRASTRUCTURE ONLY (tests/ and bench legs).

A plain-numpy restatement of select_segments / construct_segments / construct_graph
(gnn/graph.py:37-66, 68-98, 100-142) without pandas: the cross join of the hits of two layers in row
order, the wrapped dphi, phi slope and z0 cuts evaluated in the dtype of the hit columns.
PINNING: tests/test_oracle_golden.py checks it against tests/golden/segments_*.npz, which
oracle/make_golden_segments.py recorded from the reference's own construct_graph.
"""
import numpy as np


def build_segments(layer, r, phi, z, particle_id, layer_pairs, phi_slope_max, phi_slope_outer_max, z0_max,
                   outer_from_layer=5):
    """Returns (seg_start, seg_end, y): positional hit indices of the kept pairs in the reference's
    order (layer pair, hit on the first layer, hit on the second layer), and y = same particle."""
    starts, ends = [], []
    for l1, l2 in np.asarray(layer_pairs).reshape(-1, 2):
        i1 = np.nonzero(layer == l1)[0]                    # groupby('layer').get_group: row order (gnn/graph.py:79-85)
        i2 = np.nonzero(layer == l2)[0]
        if len(i1) == 0 or len(i2) == 0:
            continue                                       # the reference skips a pair with an empty layer (:86-89)
        a = np.repeat(i1, len(i2))                         # merge on evtid: left-major cross join (:53-54)
        b = np.tile(i2, len(i1))
        dphi = phi[b] - phi[a]                             # calc_dphi (:37-42)
        dphi[dphi > np.pi] -= 2 * np.pi
        dphi[dphi < -np.pi] += 2 * np.pi
        dz = z[b] - z[a]
        dr = r[b] - r[a]
        with np.errstate(divide="ignore", invalid="ignore"):
            phi_slope = dphi / dr
            z0 = z[a] - r[a] * dz / dr
        cut = phi_slope_max if l1 < outer_from_layer else phi_slope_outer_max      # hit_pairs.layer_1[0] (:65)
        keep = (np.abs(phi_slope) < cut) & (np.abs(z0) < z0_max)
        starts.append(a[keep])
        ends.append(b[keep])
    s = np.concatenate(starts) if starts else np.zeros(0, np.int64)
    e = np.concatenate(ends) if ends else np.zeros(0, np.int64)
    y = (particle_id[s] == particle_id[e]).astype(np.float32) if particle_id is not None else np.zeros(len(s), np.float32)
    return s.astype(np.int64), e.astype(np.int64), y


def features(cols, feature_scale):
    """(hits[feature_names].values / feature_scale).astype(np.float32), gnn/graph.py:118."""
    return (np.stack(cols, axis=1) / np.asarray(feature_scale)).astype(np.float32)
