"""Host-side pieces: graph tuples / NPZ format vs the reference's own files, the C host
packer vs the oracle's flattening, the module's API surface.  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_case
from oracle import segclf_oracle as O
import gnn_fpga_b200 as G
from gnn_fpga_b200 import data


def test_graph_tuples_match_reference_fields():
    assert G.Graph._fields == ("X", "Ri", "Ro", "y")
    assert G.SparseGraph._fields == ("X", "Ri_rows", "Ri_cols", "Ro_rows", "Ro_cols", "y")


def test_make_sparse_graph_equals_reference_output():
    z = np.load(os.path.join(GOLDEN, "graph_roundtrip.npz"))
    sg = G.make_sparse_graph(z["X"], z["Ri"], z["Ro"], z["y"])
    for k in ("Ri_rows", "Ri_cols", "Ro_rows", "Ro_cols"):
        assert np.array_equal(getattr(sg, k), z[k]) and getattr(sg, k).dtype == np.int64
    d = G.graph_from_sparse(sg)
    assert d.Ri.dtype == np.uint8 and np.array_equal(d.Ri, z["Ri"]) and np.array_equal(d.Ro, z["Ro"])


def test_load_file_written_by_reference_and_round_trip(tmp_path):
    """ref_saved_graph.npz was written by the reference's save_graph (oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "graph_roundtrip.npz"))
    sg = G.load_graph(os.path.join(GOLDEN, "ref_saved_graph.npz"), G.SparseGraph)
    for k in ("Ri_rows", "Ri_cols", "Ro_rows", "Ro_cols"):
        assert np.array_equal(getattr(sg, k), z[k])
    fn = str(tmp_path / "g.npz")
    G.save_graph((sg, None), fn)          # the reference's calling convention
    G.save_graph(sg, str(tmp_path / "g2.npz"))
    a, b = np.load(fn), np.load(os.path.join(GOLDEN, "ref_saved_graph.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k])
    back = G.load_graphs([fn, str(tmp_path / "g2.npz")], G.SparseGraph)
    assert all(np.array_equal(x, y) for x, y in zip(back[0], back[1]))


@pytest.mark.parametrize("gen", ["toy", "acts"])
def test_generators_emit_nonzero_order(gen):
    gs = data.toy2d_graphs(2, seed=3) if gen == "toy" else [data.acts_like_graph(25, seed=1)]
    for g in gs:
        again = G.make_sparse_graph(*G.graph_from_sparse(g))
        for a, b in zip(g, again):
            assert np.array_equal(a, b)


def test_generator_shapes():
    g = data.toy2d_graphs(1)[0]
    assert g.X.shape == (40, 3) and g.Ri_rows.shape == (144,)
    g = data.acts_like_graph(400, seed=0)
    assert g.X.shape == (4000, 3) and 18000 < g.Ri_rows.shape[0] < 22000


def test_host_packer_matches_oracle_flattening():
    gs = [data.acts_like_graph(n, seed=i) for i, n in enumerate((20, 31, 25, 8))]
    for nt in (1, 3):
        h = G.pack_sparse_batch_host(gs, n_threads=nt)
        X, src, dst, e_max = O.flatten_sparse_batch(gs)
        assert h["e_max"] == e_max and h["n_nodes"] == [g.X.shape[0] for g in gs]
        assert h["src"].dtype == torch.int32
        assert np.array_equal(h["X"].numpy(), X)
        assert np.array_equal(h["src"].numpy(), src) and np.array_equal(h["dst"].numpy(), dst)


def test_host_packer_accepts_unordered_tuples_and_int32():
    g = data.acts_like_graph(15, seed=2)
    rng = np.random.RandomState(0)
    pi, po = rng.permutation(g.Ri_rows.shape[0]), rng.permutation(g.Ro_rows.shape[0])
    shuffled = G.SparseGraph(g.X, g.Ri_rows[pi].astype(np.int32), g.Ri_cols[pi].astype(np.int32),
                             g.Ro_rows[po], g.Ro_cols[po], g.y)
    a, b = G.pack_sparse_batch_host([g]), G.pack_sparse_batch_host([shuffled])
    assert torch.equal(a["src"], b["src"]) and torch.equal(a["dst"], b["dst"])


def test_host_packer_rejects_out_of_range():
    g = data.acts_like_graph(10, seed=0)
    bad = g._replace(Ri_rows=g.Ri_rows + 1000)
    with pytest.raises(ValueError, match="out of range"):
        G.pack_sparse_batch_host([bad])
    bad = g._replace(Ro_cols=g.Ro_cols + 1)      # a column beyond len(Ri_rows): IndexError in the reference
    with pytest.raises(ValueError, match="out of range"):
        G.pack_sparse_batch_host([bad])


# ---- module surface ----------------------------------------------------------------------
def test_state_dict_keys_and_counts_match_reference():
    z = np.load(os.path.join(GOLDEN, "structure.npz"))
    for key, count in z["counts"]:
        F, h = map(int, key.split(","))
        m = G.SegmentClassifier(F, h, 1)
        assert list(m.state_dict().keys()) == list(z["keys"])
        assert sum(p.numel() for p in m.parameters()) == int(count)


def test_default_constructor_and_init_stream_match_reference():
    """Same seed -> same initial parameters as the reference (same construction order)."""
    rec = load_case("c1_toy2d_h8_it1")
    torch.manual_seed(rec["seed"])
    m = G.SegmentClassifier(rec["F"], rec["h"], rec["n_iters"])
    for k, v in m.state_dict().items():
        assert torch.equal(v, rec["params"][k]), k
    d = G.SegmentClassifier()
    assert (d.n_iters, d.input_dim, d.hidden_dim) == (3, 2, 8)


def test_reference_checkpoint_loads_and_masks_behave():
    rec = load_case("acts_masked_h8_it4")
    m = G.SegmentClassifier(rec["F"], rec["h"], rec["n_iters"], masks_e=rec["masks_e"], masks_n=rec["masks_n"])
    m.load_state_dict(rec["params"])
    e0 = m.edge_network.network[0]
    assert e0.mask_flag is True and torch.equal(e0.get_mask(), rec["masks_e"][0])
    assert "mask" not in "".join(m.state_dict().keys())          # masks are not in the state_dict
    assert torch.all(e0.weight[rec["masks_e"][0] == 0] == 0)       # set_mask zeroed the stored weights
    assert m.input_network[0].weight.shape == (8, 3)
    # attributes estimator.py:54-55 touches
    ws = [l.weight for l in m.node_network.network if hasattr(l, "weight")]
    ws += [l.weight for l in m.edge_network.network if hasattr(l, "weight")]
    assert [tuple(w.shape) for w in ws] == [(8, 33), (8, 8), (8, 22), (1, 8)]
    # model.py:100 raises TypeError for masks_n=None; the twin (and we) accept it
    G.SegmentClassifier(3, 8, 1, masks_e=None, masks_n=None)


def test_no_cpu_path_and_loud_errors():
    m = G.SegmentClassifier(3, 8, 1).eval()
    with torch.no_grad():
        with pytest.raises(G.GnnsegError, match="no CPU path"):
            m(data.toy2d_graphs(1))
        with pytest.raises(ValueError, match="CUDA tensors"):
            m([torch.zeros(1, 4, 3), torch.zeros(1, 4, 5), torch.zeros(1, 4, 5)])
    with pytest.raises(ValueError, match="Tanh"):
        G.SegmentClassifier(3, 8, 1, hidden_activation=torch.nn.ReLU)
    with pytest.raises(G.GnnsegError):
        m.edge_network(None, None, None)


def test_product_does_not_import_oracle():
    """The package must never route through oracle/ (no CPU fallback)."""
    import glob
    for fn in glob.glob(os.path.join(os.path.dirname(G.__file__), "**", "*"), recursive=True):
        if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
            assert "oracle" not in open(fn).read().replace("# oracle", ""), fn


MULTI_CASES = ["segments_multi_plain", "segments_multi_tracks", "segments_multi_nomiss"]


@pytest.mark.parametrize("name", MULTI_CASES)
def test_event_loop_and_track_selection_match_reference(name):
    """Host side of construct_graphs_device (event grouping in order of first appearance, the
    no_missing_hits / max_tracks pre-selection drawn from numpy's global generator) against the
    reference's own event loop around construct_graph (oracle/make_golden_segments.py main_multi);
    the cuts on the selected rows go through the numpy oracle: edges, labels, features bit-exact."""
    import os
    from conftest import GOLDEN
    from gnn_fpga_b200.segments import event_rows
    from oracle import segments_oracle as S
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cols = {k: z["hits_" + k] for k in ("evtid", "layer", "r", "phi", "z", "particle_id")}
    np.random.seed(int(z["seed"]))
    mt = int(z["max_tracks"])
    groups = event_rows(cols, None, None if mt < 0 else mt, bool(int(z["no_missing_hits"])))
    assert [len(g) for g in groups] == list(z["n_hits"])
    for b, g in enumerate(groups):
        s, e, y = S.build_segments(cols["layer"][g], cols["r"][g], cols["phi"][g], cols["z"][g], cols["particle_id"][g],
                                   z["layer_pairs"], float(z["phi_slope_max"]), float(z["phi_slope_outer_max"]), float(z["z0_max"]))
        assert np.array_equal(s, z["src_%d" % b]) and np.array_equal(e, z["dst_%d" % b]) and np.array_equal(y, z["y_%d" % b])
        assert np.array_equal(S.features([cols[k][g] for k in ("r", "phi", "z")], z["feature_scale"]), z["X_%d" % b])
