"""Host side of the event store (gnnseg_store_*_host, gnn_fpga_b200/store.py): layout, validation,
canonical CSR order, node renumbering, batch metadata.  Pure CPU; the device half
(gnnseg_assemble_batch) is restated in numpy here and run on the GPU in tests/test_gpu_store.py.
Replaces graph_from_sparse + merge_graphs (gnn/graph.py:28-35, gnn/trainSegmentClassifier.py:66-111)."""
import numpy as np
import pytest

from oracle import segclf_oracle as O


def assemble_numpy(store, lo, hi):
    """What gnnseg_assemble_batch computes, in numpy: (src, dst, in_ptr, in_eid, out_ptr, out_eid)."""
    meta, n, e_max, n_in, n_out = store.batch_meta(lo, hi)
    B = hi - lo
    node_off, in_base, out_base = meta[:B + 1], meta[B + 1:2 * B + 2], meta[2 * B + 2:]
    X, in_ptr_l, out_ptr_l, in_col, out_col = (t.numpy() for t in store.slices(lo, hi))
    if store.col_bytes == 2:
        in_col, out_col = in_col.view(np.uint16), out_col.view(np.uint16)
    src = np.full(B * e_max, -1, np.int64)
    dst = np.full(B * e_max, -1, np.int64)
    in_ptr, out_ptr = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
    in_eid, out_eid = np.zeros(n_in, np.int64), np.zeros(n_out, np.int64)
    for b in range(B):
        for i in range(node_off[b + 1] - node_off[b]):
            g = node_off[b] + i
            lp = g + b
            for ptr_l, base, col, key, ptr, eid in ((in_ptr_l, in_base, in_col, dst, in_ptr, in_eid),
                                                    (out_ptr_l, out_base, out_col, src, out_ptr, out_eid)):
                beg, end = base[b] + ptr_l[lp], base[b] + ptr_l[lp + 1]
                ptr[g], ptr[g + 1] = beg, end
                for k in range(beg, end):
                    slot = b * e_max + int(col[k])
                    assert key[slot] == -1
                    key[slot] = g
                    eid[k] = slot
    return X, src, dst, in_ptr, in_eid, out_ptr, out_eid, e_max


def small_graphs():
    from gnn_fpga_b200 import data
    return [data.acts_like_graph(n, seed=i) for i, n in enumerate((30, 44, 25))] + data.toy2d_graphs(2)


def test_store_without_reorder_is_the_np_nonzero_csr():
    """reorder off: the assembled batch is bit-identical to flatten_sparse_batch + np.nonzero-order CSRs."""
    from gnn_fpga_b200.store import GraphStore
    graphs = small_graphs()
    store = GraphStore.from_sparse_graphs(graphs, reorder=False, pin=False)
    assert len(store) == len(graphs) and store.order_column == -1 and store.col_bytes == 2
    for lo, hi in ((0, 5), (1, 3), (4, 5)):
        X, src, dst, in_ptr, in_eid, out_ptr, out_eid, e_max = assemble_numpy(store, lo, hi)
        Xr, sr, dr, er = O.flatten_sparse_batch(graphs[lo:hi])
        assert e_max == er and np.array_equal(X, Xr) and np.array_equal(src, sr) and np.array_equal(dst, dr)
        for key, ptr, eid in ((dr, in_ptr, in_eid), (sr, out_ptr, out_eid)):
            p, e = O.csr_from_keys(key, Xr.shape[0])
            assert np.array_equal(ptr, p) and np.array_equal(eid, e)
    y = store.targets(0, 5).numpy()
    for b, g in enumerate(graphs):
        assert np.array_equal(y[b, :g.y.shape[0]], g.y) and not y[b, g.y.shape[0]:].any()


def test_store_reorder_keeps_the_graph():
    """reorder on: nodes are renumbered along the most edge-local feature (phi for the ACTS-like events);
    undoing the permutation gives back every endpoint, and the CSR rows stay in ascending slot order."""
    from gnn_fpga_b200 import data
    from gnn_fpga_b200.store import GraphStore
    graphs = [data.acts_like_graph(400, seed=s) for s in range(3)]
    store = GraphStore.from_sparse_graphs(graphs, reorder=True, pin=False)
    assert store.order_column == 1                         # X = (r, phi, z): edges are local in phi
    X, src, dst, in_ptr, in_eid, out_ptr, out_eid, e_max = assemble_numpy(store, 0, 3)
    Xr, sr, dr, er = O.flatten_sparse_batch(graphs)
    perm = store.perm.numpy().astype(np.int64) + np.repeat(store.node_off[:-1], np.diff(store.node_off))   # internal -> original
    assert sorted(perm.tolist()) == list(range(Xr.shape[0]))
    assert np.array_equal(X, Xr[perm])
    for a, ref in ((src, sr), (dst, dr)):
        real = ref >= 0
        assert np.array_equal(a >= 0, real) and np.array_equal(perm[a[real]], ref[real])
    for b in range(3):                                     # sorted along phi inside every event
        n0, n1 = int(store.node_off[b]), int(store.node_off[b + 1])
        assert np.all(np.diff(X[n0:n1, 1]) >= 0)
    for key, ptr, eid in ((dst, in_ptr, in_eid), (src, out_ptr, out_eid)):
        p, e = O.csr_from_keys(key, Xr.shape[0])
        assert np.array_equal(ptr, p) and np.array_equal(eid, e)
    # auto: small events are left alone
    assert GraphStore.from_sparse_graphs(graphs, reorder="auto", pin=False).order_column == 1
    assert GraphStore.from_sparse_graphs([data.acts_like_graph(60, seed=0)], reorder="auto", pin=False).order_column == -1


def test_store_accepts_unsorted_tuples_and_int32_columns():
    """Index arrays in any order give the same store as np.nonzero order; events with more than 65536
    edges switch the column arrays to int32."""
    from gnn_fpga_b200 import SparseGraph, data
    from gnn_fpga_b200.store import GraphStore
    g = data.acts_like_graph(50, seed=7)
    rng = np.random.RandomState(0)
    pi, po = rng.permutation(g.Ri_rows.shape[0]), rng.permutation(g.Ro_rows.shape[0])
    shuffled = SparseGraph(g.X, g.Ri_rows[pi], g.Ri_cols[pi], g.Ro_rows[po], g.Ro_cols[po], g.y)
    a = GraphStore.from_sparse_graphs([g], reorder=False, pin=False)
    b = GraphStore.from_sparse_graphs([shuffled], reorder=False, pin=False)
    for ta, tb in zip(a.slices(0, 1), b.slices(0, 1)):      # X, both row pointers, both column arrays
        assert np.array_equal(ta.numpy(), tb.numpy())
    n = 70000
    big = SparseGraph(np.zeros((n + 1, 3), np.float32), np.arange(1, n + 1), np.arange(n), np.arange(n), np.arange(n),
                      np.zeros(n, np.float32))
    s = GraphStore.from_sparse_graphs([big], reorder=False, pin=False)
    assert s.col_bytes == 4
    X, src, dst, in_ptr, in_eid, out_ptr, out_eid, e_max = assemble_numpy(s, 0, 1)
    assert e_max == n and np.array_equal(src, np.arange(n)) and np.array_equal(dst, np.arange(1, n + 1))


def test_store_rejects_bad_graphs():
    """Out-of-range indices and a column listed twice (a hyper-edge: the dense reference would sum two
    rows, gnn/model.py:71-72) are ValueErrors naming the event, as for the dense entry point."""
    from gnn_fpga_b200 import SparseGraph, data
    from gnn_fpga_b200.store import GraphStore
    good = data.acts_like_graph(20, seed=1)
    rows = good.Ri_rows.copy(); rows[3] = good.X.shape[0]
    with pytest.raises(ValueError, match="graph 1.*out of range"):
        GraphStore.from_sparse_graphs([good, good._replace(Ri_rows=rows)], pin=False)
    cols = good.Ro_cols.copy(); cols[5] = cols[4]
    with pytest.raises(ValueError, match="graph 0.*more than one"):
        GraphStore.from_sparse_graphs([good._replace(Ro_cols=cols), good], pin=False)
    cols = good.Ri_cols.copy(); cols[0] = good.Ri_rows.shape[0] + good.Ro_rows.shape[0]
    with pytest.raises(ValueError, match="out of range"):
        GraphStore.from_sparse_graphs([good._replace(Ri_cols=cols)], pin=False)
    # a column listed only in Ro (an edge without an end node) is a legal graph: it keeps its column
    half = good._replace(Ri_rows=good.Ri_rows[good.Ri_cols != good.Ri_cols.max()], Ri_cols=good.Ri_cols[good.Ri_cols != good.Ri_cols.max()])
    st = GraphStore.from_sparse_graphs([half], reorder=False, pin=False)
    assert int(st.n_in[0]) == good.Ri_rows.shape[0] and st.batch_meta(0, 1)[2] == good.Ri_rows.shape[0]


def test_batches_follow_the_reference_generator():
    """store.batches cuts graphs[j : j + batch_size] like batch_generator (gnn/trainSegmentClassifier.py:99-103)."""
    from gnn_fpga_b200.store import GraphStore
    store = GraphStore.from_sparse_graphs(small_graphs(), reorder=False, pin=False)
    assert [(b.lo, b.hi) for b in store.batches(2)] == [(0, 2), (2, 4), (4, 5)]
    assert [(b.lo, b.hi) for b in store.batches(2, n_samples=4)] == [(0, 2), (2, 4)]
    assert store.max_batch_shape(store.batches(2))[4] == 2
    assert store.h2d_bytes(0, 5) == sum(t.numel() * t.element_size() for t in store.slices(0, 5)) + 12 * 6
