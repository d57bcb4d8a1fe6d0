"""Golden vectors for GPU segment construction: the reference's own construct_graph
(gnn/graph.py:100-142, select_segments :44-66) run HERE on synthetic hit DataFrames.
Writes tests/golden/segments_*.npz.  TEST INFRASTRUCTURE ONLY (needs /root/reference, pandas)."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/gnn")
sys.dont_write_bytecode = True
import graph as ref_graph                                   # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden")
BARREL_R = np.array([32., 72., 116., 172., 260., 360., 500., 660., 820., 1020.])


def synth_hits(n_tracks, seed, dtype, shuffle):
    """ACTS-like barrel hits as the reference's pipelines hold them: one row per hit with evtid,
    layer, r, phi, z, particle_id; row order = file order (shuffled here: NOT grouped by layer)."""
    rng = np.random.RandomState(seed)
    L = len(BARREL_R)
    phi0 = rng.uniform(-np.pi, np.pi, n_tracks)
    kappa = rng.normal(0.0, 2.5e-4, n_tracks)
    z0 = rng.normal(0.0, 50.0, n_tracks)
    cot = rng.uniform(-1.0, 1.0, n_tracks)
    layer = np.repeat(np.arange(L), n_tracks)
    track = np.tile(np.arange(n_tracks), L)
    r = BARREL_R[layer] + rng.normal(0, 0.5, layer.shape[0])
    phi = (phi0[track] + kappa[track] * r + np.pi) % (2 * np.pi) - np.pi
    z = z0[track] + cot[track] * r
    df = pd.DataFrame({"evtid": np.zeros(layer.shape[0], np.int64), "layer": layer.astype(np.int64),
                       "r": r.astype(dtype), "phi": phi.astype(dtype), "z": z.astype(dtype),
                       "particle_id": (track + 1000).astype(np.int64)})
    keep = rng.uniform(size=len(df)) > 0.08                 # missing hits
    df = df[keep]
    if shuffle:
        df = df.sample(frac=1.0, random_state=seed)
    return df                                               # keeps the (non-contiguous) original index labels


def main():
    l = np.arange(10)
    layer_pairs = np.stack([l[:-1], l[1:]], axis=1)
    cases = [("f32_small", 30, 1, np.float32, True, 0.003, 0.006, 200.0),
             ("f32_event", 400, 2, np.float32, True, 0.0006, 0.0012, 200.0),
             ("f64_small", 40, 3, np.float64, False, 0.002, 0.004, 150.0),
             ("f32_wide", 25, 4, np.float32, True, 0.05, 0.05, 2000.0)]       # wide cuts: wrap-around pairs pass
    for name, n_tracks, seed, dtype, shuffle, c_in, c_out, z0_max in cases:
        hits = synth_hits(n_tracks, seed, dtype, shuffle)
        feature_names = ["r", "phi", "z"]
        feature_scale = np.array([1000., np.pi / 8, 1000.])
        sg, segments = ref_graph.construct_graph(hits, layer_pairs, c_in, c_in, c_out, z0_max, feature_names, feature_scale)
        n_edges = sg.y.shape[0]
        # positional endpoints per edge, as construct_graph computes them (gnn/graph.py:128-130)
        hit_idx = pd.Series(np.arange(len(hits)), index=hits.index)
        seg_start = hit_idx.loc[segments.index_1].values
        seg_end = hit_idx.loc[segments.index_2].values
        assert np.array_equal(sg.Ro_rows[np.argsort(sg.Ro_cols, kind="stable")], seg_start)
        np.savez_compressed(os.path.join(OUT, "segments_%s.npz" % name),
                            layer=hits.layer.values.astype(np.int32), r=hits.r.values, phi=hits.phi.values, z=hits.z.values,
                            particle_id=hits.particle_id.values, layer_pairs=layer_pairs.astype(np.int32),
                            phi_slope_max=c_in, phi_slope_outer_max=c_out, z0_max=z0_max, feature_scale=feature_scale,
                            seg_start=seg_start.astype(np.int64), seg_end=seg_end.astype(np.int64), y=sg.y, X=sg.X,
                            Ri_rows=sg.Ri_rows, Ri_cols=sg.Ri_cols, Ro_rows=sg.Ro_rows, Ro_cols=sg.Ro_cols)
        print("segments_%-10s hits=%5d edges=%6d true=%d dtype=%s" % (name, len(hits), n_edges, int(sg.y.sum()), np.dtype(dtype).name))


def main_multi():
    """Several events in one hit table, rows of different events interleaved: the reference's event loop
    (construct_graphs, gnn/graph.py:152-174: groupby('evtid'), events in order of first appearance)
    around its construct_graph, with the max_tracks / no_missing_hits pre-selection (gnn/graph.py:106-114).
    np.random is seeded once before the loop; the drop-in must draw the same samples."""
    l = np.arange(10)
    layer_pairs = np.stack([l[:-1], l[1:]], axis=1)
    feature_names = ["r", "phi", "z"]
    feature_scale = np.array([1000., np.pi / 8, 1000.])
    for name, sizes, max_tracks, no_missing, seed in (("multi_plain", (30, 18, 41), None, False, 5),
                                                     ("multi_tracks", (40, 25, 33), 15, False, 6),
                                                     ("multi_nomiss", (30, 36), 20, True, 7)):
        parts = []
        for e, n in enumerate(sizes):
            h = synth_hits(n, 100 * seed + e, np.float32, False)
            h["evtid"] = 7 + 3 * e
            h["particle_id"] += 100000 * e
            parts.append(h)
        hits = pd.concat(parts, ignore_index=True).sample(frac=1.0, random_state=seed)
        c_in, c_out, z0_max = 0.003, 0.006, 200.0
        np.random.seed(40 + seed)
        # Work-around needed to run the reference here: pandas 3 hands out read-only `.values`, and
        # gnn/graph.py:112 shuffles that array in place (ValueError).  The stand-in re-enables writing on
        # the very same array and calls numpy's shuffle: same generator, same draws, same in-place result.
        orig_shuffle = np.random.shuffle

        def writable_shuffle(a):
            a.flags.writeable = True
            orig_shuffle(a)
        np.random.shuffle = writable_shuffle
        groups = hits.groupby("evtid")
        rec = {"seed": 40 + seed, "max_tracks": -1 if max_tracks is None else max_tracks, "no_missing_hits": int(no_missing),
               "layer_pairs": layer_pairs.astype(np.int32), "phi_slope_max": c_in, "phi_slope_outer_max": c_out,
               "z0_max": z0_max, "feature_scale": feature_scale}
        for k in ("evtid", "layer", "r", "phi", "z", "particle_id"):
            rec["hits_" + k] = hits[k].values
        n_hits, n_edges = [], []
        for b, evtid in enumerate(hits.evtid.unique()):
            evt_hits = groups.get_group(evtid)
            sg, segments = ref_graph.construct_graph(evt_hits, layer_pairs, c_in, c_in, c_out, z0_max, feature_names,
                                                     feature_scale, max_tracks=max_tracks, no_missing_hits=no_missing)
            src = sg.Ro_rows[np.argsort(sg.Ro_cols, kind="stable")]
            dst = sg.Ri_rows[np.argsort(sg.Ri_cols, kind="stable")]
            rec["src_%d" % b], rec["dst_%d" % b], rec["y_%d" % b], rec["X_%d" % b] = src, dst, sg.y, sg.X
            n_hits.append(sg.X.shape[0])
            n_edges.append(sg.y.shape[0])
        np.random.shuffle = orig_shuffle
        rec["n_hits"], rec["n_edges"] = np.array(n_hits), np.array(n_edges)
        np.savez_compressed(os.path.join(OUT, "segments_%s.npz" % name), **rec)
        print("segments_%-12s events=%d hits=%s edges=%s" % (name, len(n_hits), n_hits, n_edges))


if __name__ == "__main__":
    main()
    main_multi()
