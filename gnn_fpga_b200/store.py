"""GraphStore: a data set of graphs laid out once, at load time, for the GPU.

The reference keeps `graphs = load_graphs(filenames, SparseGraph)` (gnn/trainSegmentClassifier.py:129)
and, for EVERY batch, densifies a slice of that list on the CPU (graph_from_sparse + merge_graphs +
np_to_torch, gnn/trainSegmentClassifier.py:97-111).  Here the list is narrowed, validated and packed
ONCE into one pinned host arena (gnnseg_store_fill_host); a batch of consecutive events then costs the
host nothing but five contiguous asynchronous copies, and the device builds the batch graph
(gnnseg_assemble_batch).  `store.graphs` keeps the original tuples, so code written against the
reference's list keeps working.

    store = GraphStore.from_sparse_graphs(graphs)          # or GraphStore.from_files(filenames)
    for scores in model.predict_stream(store.batches(64)): ...
    gen = batch_generator(store, n_samples, batch_size)    # gnn_fpga_b200.loader: (inputs, target) pairs
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

REORDER = {False: 0, None: 0, "none": 0, True: 1, "always": 1, "auto": 2}


class StoreBatch:
    """Events [lo, hi) of a GraphStore: what SegmentClassifier.forward / predict_stream accept in place
    of a list of SparseGraph tuples."""
    __slots__ = ("store", "lo", "hi")

    def __init__(self, store, lo, hi):
        if not (0 <= lo < hi <= len(store)):
            raise IndexError("batch [%d, %d) outside a store of %d events" % (lo, hi, len(store)))
        self.store, self.lo, self.hi = store, int(lo), int(hi)

    def __len__(self):
        return self.hi - self.lo

    @property
    def graphs(self):
        return self.store.graphs[self.lo:self.hi]


class GraphStore:
    """Pinned host arena of a list of graphs (include/gnnseg.h, GnnsegStoreLayout)."""

    def __init__(self, graphs, arena, layout, order_column, n_edges=None):
        self.graphs = graphs                      # the original SparseGraph tuples (numpy views), as load_graphs gives them
        self.arena = arena                        # uint8 tensor, pinned when CUDA is available
        self.layout = layout
        self.order_column = order_column          # feature column the nodes were renumbered along, or -1
        L = layout
        B = int(L.n_events)
        self.F = int(L.n_features)
        self.col_bytes = int(L.col_bytes)

        def i64(off):
            return arena[off:off + 8 * (B + 1)].view(torch.int64).numpy()

        self.node_off, self.in_off, self.out_off, self.y_off = i64(L.o_node_off), i64(L.o_in_off), i64(L.o_out_off), i64(L.o_y_off)
        tn, ti, to, ty = int(L.total_nodes), int(L.total_in), int(L.total_out), int(L.total_y)
        col_dt = torch.int16 if self.col_bytes == 2 else torch.int32     # uint16 bit patterns travel as int16
        self.X = arena[L.o_X:L.o_X + 4 * tn * self.F].view(torch.float32).view(tn, self.F)
        self.in_ptr = arena[L.o_in_ptr:L.o_in_ptr + 4 * (tn + B)].view(torch.int32)
        self.out_ptr = arena[L.o_out_ptr:L.o_out_ptr + 4 * (tn + B)].view(torch.int32)
        self.in_col = arena[L.o_in_col:L.o_in_col + self.col_bytes * ti].view(col_dt)
        self.out_col = arena[L.o_out_col:L.o_out_col + self.col_bytes * to].view(col_dt)
        self.y = arena[L.o_y:L.o_y + 4 * ty].view(torch.float32)
        self.perm = arena[L.o_perm:L.o_perm + 4 * tn].view(torch.int32)
        # edge columns per event: len(Ri_rows), as graph_from_sparse counts them (gnn/graph.py:30), or more when
        # the tuple has columns without an end node
        self.n_in = np.diff(self.in_off) if n_edges is None else np.asarray(n_edges, dtype=np.int64)

    def __len__(self):
        return int(self.layout.n_events)

    # -- construction ------------------------------------------------------------------
    @classmethod
    def from_sparse_graphs(cls, graphs, reorder="auto", n_threads=0, pin=None, arena=None):
        """List of host SparseGraph tuples (gnn/graph.py:20-21) -> store.  Raises ValueError for an index
        out of range or a column listed twice in Ri / Ro (the dense entry point's BAD_HYPEREDGE).
        `arena`: a uint8 tensor to fill instead of a new allocation when it is large enough (per-batch
        stores of a pipeline reuse theirs: allocating pinned memory costs more than filling it)."""
        L = _lib.lib()
        graphs = list(graphs)
        B = len(graphs)
        keep = []

        def arr(a, dt):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
                a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        F = int(graphs[0].X.shape[1]) if B else 1
        Xs = [arr(g.X, np.float32) for g in graphs]
        cols = [[arr(g[k], np.int64) for g in graphs] for k in (1, 2, 3, 4)]
        ys = [arr(np.asarray(g.y).reshape(-1), np.float32) for g in graphs]
        for b in range(B):
            if Xs[b].ndim != 2 or Xs[b].shape[1] != F:
                raise ValueError("graph %d: X must be (N, %d)" % (b, F))
            if cols[1][b].shape[0] != cols[0][b].shape[0] or cols[3][b].shape[0] != cols[2][b].shape[0]:
                raise ValueError("graph %d: rows/cols length mismatch" % b)
        n_nodes = np.fromiter((x.shape[0] for x in Xs), dtype=np.int64, count=B)
        n_in = np.fromiter((a.shape[0] for a in cols[0]), dtype=np.int64, count=B)
        n_out = np.fromiter((a.shape[0] for a in cols[2]), dtype=np.int64, count=B)
        n_y = np.fromiter((a.shape[0] for a in ys), dtype=np.int64, count=B)
        # edge columns of an event: len(Ri_rows) as graph_from_sparse has it; a tuple that lists a column only in
        # Ro (an edge without an end node, which only a dense graph can express in the reference) keeps that column
        # (every column is listed in Ri or in Ro, so there are at most len(Ri_rows) + len(Ro_rows) of them: a larger
        # column id is an error, reported by the fill as out of range)
        n_edges = np.fromiter((max(int(n_in[b]), min(max(int(cols[1][b].max(initial=-1)), int(cols[3][b].max(initial=-1))) + 1,
                                                     int(n_in[b] + n_out[b])))
                               for b in range(B)), dtype=np.int64, count=B)
        ptr = lambda a: a.__array_interface__["data"][0]
        layout = _lib.GnnsegStoreLayout()
        _lib.check(L.gnnseg_store_plan_host(B, F, ptr(n_nodes), ptr(n_in), ptr(n_out), ptr(n_y), ptr(n_edges), C.byref(layout)),
                   "gnnseg_store_plan_host")
        if pin is None:
            pin = torch.cuda.is_available()
        if arena is None or arena.numel() < int(layout.bytes):
            arena = torch.empty(max(int(layout.bytes * (1.25 if arena is not None else 1.0)), 1), dtype=torch.uint8, pin_memory=pin)

        def pp(arrs):
            t = np.fromiter((ptr(a) for a in arrs), dtype=np.uintp, count=B)
            keep.append(t)
            return ptr(t)

        info = np.full(2, -1, dtype=np.int32)
        rc = L.gnnseg_store_fill_host(C.byref(layout), pp(Xs), ptr(n_nodes), pp(cols[0]), pp(cols[1]), pp(cols[2]), pp(cols[3]),
                                      ptr(n_in), ptr(n_out), pp(ys), ptr(n_y), ptr(n_edges), REORDER[reorder], n_threads,
                                      arena.data_ptr(), ptr(info))
        if rc == _lib.EHYPEREDGE:
            raise ValueError("graph %d: a column of Ri or Ro has more than one non-zero entry" % int(info[0]))
        if rc == -1:
            raise ValueError("graph %d: SparseGraph index out of range (row >= n_nodes or col >= len(Ri_rows))" % int(info[0]))
        _lib.check(rc, "gnnseg_store_fill_host")
        return cls(graphs, arena, layout, int(info[1]), n_edges)

    @classmethod
    def from_files(cls, filenames, reorder="auto", n_threads=0):
        """Graph files in the reference's NPZ format (gnn/graph.py:179-194), memory mapped and parsed by the
        library (no np.load); `store.graphs` are zero-copy views of the mappings."""
        from .graph import load_graphs_mapped
        return cls.from_sparse_graphs(load_graphs_mapped([os.fspath(f) for f in filenames]), reorder=reorder, n_threads=n_threads)

    # -- batches -----------------------------------------------------------------------
    def batch(self, lo, hi):
        return StoreBatch(self, lo, hi)

    def batches(self, batch_size, n_samples=None):
        """Consecutive batches as the reference's batch_generator cuts them
        (gnn/trainSegmentClassifier.py:99-103: batch j = graphs[j : j + batch_size])."""
        n = len(self) if n_samples is None else min(int(n_samples), len(self))
        return [StoreBatch(self, j, min(j + batch_size, n)) for j in range(0, n, batch_size)]

    def batch_meta(self, lo, hi):
        """(meta int32 [3*(B+1)], n_nodes, e_max, n_in, n_out) of events [lo, hi): the small per-batch
        record gnnseg_assemble_batch wants, relative to the batch."""
        B = hi - lo
        meta = np.empty(3 * (B + 1), dtype=np.int32)
        meta[0:B + 1] = self.node_off[lo:hi + 1] - self.node_off[lo]
        meta[B + 1:2 * B + 2] = self.in_off[lo:hi + 1] - self.in_off[lo]
        meta[2 * B + 2:] = self.out_off[lo:hi + 1] - self.out_off[lo]
        return meta, int(meta[B]), int(self.n_in[lo:hi].max()) if B else 0, int(meta[2 * B + 1]), int(meta[3 * B + 2])

    def slices(self, lo, hi):
        """The five contiguous host slices of events [lo, hi) (pinned views, no copy)."""
        n0, n1 = int(self.node_off[lo]), int(self.node_off[hi])
        i0, i1 = int(self.in_off[lo]), int(self.in_off[hi])
        o0, o1 = int(self.out_off[lo]), int(self.out_off[hi])
        return (self.X[n0:n1], self.in_ptr[n0 + lo:n1 + hi], self.out_ptr[n0 + lo:n1 + hi], self.in_col[i0:i1], self.out_col[o0:o1])

    def h2d_bytes(self, lo, hi):
        return sum(t.numel() * t.element_size() for t in self.slices(lo, hi)) + 12 * (hi - lo + 1)

    def targets(self, lo, hi, out=None):
        """Padded (B, e_max) float32 labels of events [lo, hi), zero where an event has fewer edges:
        merge_graphs' batch_y (gnn/trainSegmentClassifier.py:82-92)."""
        B = hi - lo
        e_max = int(self.n_in[lo:hi].max()) if B else 0
        if out is None:
            out = torch.zeros((B, e_max), dtype=torch.float32, pin_memory=self.arena.is_pinned())
        else:
            out.zero_()
        for b in range(B):
            y0, y1 = int(self.y_off[lo + b]), int(self.y_off[lo + b + 1])
            k = min(y1 - y0, e_max)
            out[b, :k] = self.y[y0:y0 + k]
        return out

    def max_batch_shape(self, batches):
        """Largest (n_nodes, n_in, n_out, n_slots, B) over `batches`: sizes the reusable device buffers."""
        mx = [0, 0, 0, 0, 0]
        for sb in batches:
            _, n, e_max, ni, no = self.batch_meta(sb.lo, sb.hi)
            B = sb.hi - sb.lo
            for k, v in enumerate((n, ni, no, B * e_max, B)):
                mx[k] = max(mx[k], v)
        return tuple(mx)
