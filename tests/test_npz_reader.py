"""The memory-mapped NPZ graph reader (gnnseg_npz_open_graph_host, replaces load_graph,
gnn/graph.py:188-191) against np.load on files in the reference's format, including the file the
reference's own save_graph wrote (tests/golden/ref_saved_graph.npz).  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gnn_fpga_b200 import (NpzGraphFile, SparseGraph, _lib, data, load_graph, load_graphs, load_graphs_mapped,
                           pack_npz_batch_host, pack_sparse_batch_host, save_graph, save_graphs)


def _same(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


def test_reads_the_reference_writers_file():
    path = os.path.join(GOLDEN, "ref_saved_graph.npz")
    ref = load_graph(path, SparseGraph)                      # np.load, as the reference reads it
    f = NpzGraphFile(path)
    for a, b in zip(f.graph, ref):
        _same(a, b)
    assert not f.graph.X.flags.writeable
    f.close()
    assert f.graph is None


def test_many_files_round_trip_and_pack_identically(tmp_path):
    graphs = [data.acts_like_graph(n, seed=i) for i, n in enumerate((17, 40, 3, 25))]
    names = [str(tmp_path / ("graph%06i.npz" % i)) for i in range(len(graphs))]     # prepareGraphs.py:235 naming
    save_graphs(graphs, names)
    mapped = load_graphs_mapped(names)
    loaded = load_graphs(names, SparseGraph)
    for g, h in zip(mapped, loaded):
        for a, b in zip(g, h):
            _same(a, b)
    a = pack_sparse_batch_host(list(mapped), n_threads=2)
    b = pack_sparse_batch_host(loaded, n_threads=2)
    assert a["e_max"] == b["e_max"] and a["n_nodes"] == b["n_nodes"]
    c = pack_npz_batch_host(names, n_threads=2)              # files -> packed batch, all in the library
    assert c["e_max"] == b["e_max"] and c["n_nodes"] == b["n_nodes"]
    for k in ("X", "src", "dst"):
        assert np.array_equal(a[k].numpy(), b[k].numpy()) and np.array_equal(c[k].numpy(), b[k].numpy())
    bad = names[:2] + [str(tmp_path / "absent.npz")]
    with pytest.raises(_lib.GnnsegError, match="EIO"):
        pack_npz_batch_host(bad)


def test_empty_graph_and_missing_y(tmp_path):
    e = np.zeros(0, np.int64)
    p1 = str(tmp_path / "empty.npz")
    save_graph(SparseGraph(np.zeros((0, 3), np.float32), e, e, e, e, np.zeros(0, np.float32)), p1)
    g = NpzGraphFile(p1).graph
    assert g.X.shape == (0, 3) and g.Ri_rows.shape == (0,) and g.y.shape == (0,)
    p2 = str(tmp_path / "no_y.npz")
    sg = data.acts_like_graph(9, seed=1)
    np.savez(p2, **{k: v for k, v in sg._asdict().items() if k != "y"})
    f = NpzGraphFile(p2)
    _same(f.graph.Ri_cols, sg.Ri_cols)
    assert f.graph.y.shape == (0,)


def test_misaligned_member_is_copied(tmp_path):
    """An extra 1-byte member first shifts every later member off its natural alignment."""
    sg = data.acts_like_graph(12, seed=2)
    p = str(tmp_path / "shifted.npz")
    np.savez(p, a=np.zeros(1, np.uint8), **sg._asdict())
    f = NpzGraphFile(p)
    for a, b in zip(f.graph, sg):
        _same(a, b)


def test_rejects_what_it_cannot_read(tmp_path):
    sg = data.acts_like_graph(9, seed=3)
    L = _lib.lib()
    out = _lib.GnnsegNpzGraph()
    # compressed archive
    p = str(tmp_path / "compressed.npz")
    np.savez_compressed(p, **sg._asdict())
    assert L.gnnseg_npz_open_graph_host(p.encode(), C.byref(out)) == -2 and not out.handle
    with pytest.raises(_lib.GnnsegError, match="unsupported"):
        NpzGraphFile(p)
    # wrong dtype (float64 X), a member missing, not a zip, no file
    p = str(tmp_path / "f64.npz")
    np.savez(p, **dict(sg._asdict(), X=sg.X.astype(np.float64)))
    assert L.gnnseg_npz_open_graph_host(p.encode(), C.byref(out)) == -2
    p = str(tmp_path / "missing.npz")
    np.savez(p, X=sg.X, Ri_rows=sg.Ri_rows)
    assert L.gnnseg_npz_open_graph_host(p.encode(), C.byref(out)) == -7
    p = str(tmp_path / "garbage.npz")
    open(p, "wb").write(b"not a zip archive at all, but longer than twenty-two bytes")
    assert L.gnnseg_npz_open_graph_host(p.encode(), C.byref(out)) == -7
    assert L.gnnseg_npz_open_graph_host(str(tmp_path / "absent.npz").encode(), C.byref(out)) == -6
    assert L.gnnseg_npz_open_graph_host(None, C.byref(out)) == -1
    # rows / cols of different lengths
    p = str(tmp_path / "ragged.npz")
    np.savez(p, **dict(sg._asdict(), Ri_cols=sg.Ri_cols[:-1]))
    assert L.gnnseg_npz_open_graph_host(p.encode(), C.byref(out)) == -7
