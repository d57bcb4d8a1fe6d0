"""Graph tuples and the device batch.

Host side: the reference's graph.py surface (gnn/graph.py:18-35,179-194) with the same
names, field order, dtypes and NPZ keys, so files and tuples are interchangeable.

Device side: `DeviceGraphBatch`, the flattened block-diagonal batch the CUDA path works on
(include/gnnseg.h, GnnsegGraph).  It is built either from the dense (X, Ri, Ro) tensors the
reference's batch_generator yields (gnn/trainSegmentClassifier.py:97-111) or straight from a
list of SparseGraph tuples without ever densifying.
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _lib

# gnn/graph.py:18-21
Graph = namedtuple("Graph", ["X", "Ri", "Ro", "y"])
SparseGraph = namedtuple("SparseGraph", ["X", "Ri_rows", "Ri_cols", "Ro_rows", "Ro_cols", "y"])


def make_sparse_graph(X, Ri, Ro, y):
    """Dense incidence matrices -> index tuples, in np.nonzero (row-major) order
    (gnn/graph.py:23-26).  That order is destination-CSR for Ri and source-CSR for Ro."""
    in_rows, in_cols = np.nonzero(Ri)
    out_rows, out_cols = np.nonzero(Ro)
    return SparseGraph(X, in_rows, in_cols, out_rows, out_cols, y)


def graph_from_sparse(sparse_graph, dtype=np.uint8):
    """Index tuples -> dense (N, E) incidence matrices (gnn/graph.py:28-35).  The edge count
    is len(Ri_rows), as in the reference.  Host-only helper for small graphs: the device path
    never calls it."""
    n_nodes = sparse_graph.X.shape[0]
    n_edges = sparse_graph.Ri_rows.shape[0]
    dense = [np.zeros((n_nodes, n_edges), dtype=dtype) for _ in range(2)]
    dense[0][sparse_graph.Ri_rows, sparse_graph.Ri_cols] = 1
    dense[1][sparse_graph.Ro_rows, sparse_graph.Ro_cols] = 1
    return Graph(sparse_graph.X, dense[0], dense[1], sparse_graph.y)


def save_graph(graph, filename):
    """NPZ with the six SparseGraph keys.  Like the reference (gnn/graph.py:179-181) this
    accepts the `(SparseGraph, segments)` pair construct_graph returns; a bare SparseGraph
    is accepted too."""
    sg = graph if isinstance(graph, SparseGraph) else graph[0]
    np.savez(filename, **sg._asdict())


def save_graphs(graphs, filenames):
    for graph, filename in zip(graphs, filenames):
        save_graph(graph, filename)


def load_graph(filename, graph_type=Graph):
    """gnn/graph.py:188-191."""
    with np.load(filename) as f:
        return graph_type(**{k: f[k] for k in f.files})


def load_graphs(filenames, graph_type=Graph):
    return [load_graph(f, graph_type) for f in filenames]


class _NpzMapping:
    """Owns one gnnseg_npz_open_graph_host mapping; unmapped when the last array viewing it dies."""

    def __init__(self, filename):
        import ctypes as C
        self.g = _lib.GnnsegNpzGraph()
        rc = _lib.lib().gnnseg_npz_open_graph_host(str(filename).encode(), C.byref(self.g))
        if rc != 0:
            self.g = None
            _lib.check(rc, "gnnseg_npz_open_graph_host(%s)" % filename)

    def __del__(self):
        try:
            if self.g is not None:
                import ctypes as C
                _lib.lib().gnnseg_npz_close_graph_host(C.byref(self.g))
                self.g = None
        except Exception:
            pass


class NpzGraphFile:
    """One graph file of the reference's on-disk format (gnn/graph.py:179-194), memory mapped by the
    library (gnnseg_npz_open_graph_host).  `.graph` is a SparseGraph whose arrays are zero-copy,
    read-only views of the mapping; every array keeps the mapping alive, so the tuple (or any slice of
    its arrays) may outlive this object.  They go straight into pack_sparse_batch_host /
    SegmentClassifier(...)."""

    def __init__(self, filename):
        import ctypes as C
        m = _NpzMapping(filename)
        g = m.g

        def view(ptr, ctype, shape):
            n = int(np.prod(shape))
            if n == 0 or not ptr:
                return np.zeros(shape, dtype=np.dtype(ctype))
            buf = (ctype * n).from_address(ptr)
            buf._gnnseg_mapping = m                       # the array's base chain holds the mapping
            a = np.ctypeslib.as_array(buf).reshape(shape)
            a.flags.writeable = False
            return a

        self.graph = SparseGraph(
            view(g.X, C.c_float, (g.n_nodes, g.n_features)),
            view(g.Ri_rows, C.c_int64, (g.n_in,)), view(g.Ri_cols, C.c_int64, (g.n_in,)),
            view(g.Ro_rows, C.c_int64, (g.n_out,)), view(g.Ro_cols, C.c_int64, (g.n_out,)),
            view(g.y, C.c_float, (g.n_y,)))

    def close(self):
        """Drop this object's views (the file is unmapped once no array refers to it any more)."""
        self.graph = None


def load_graphs_mapped(filenames):
    """load_graphs(filenames, SparseGraph) (gnn/graph.py:193-194) without np.load: every file is
    memory mapped and parsed by the library; the result is a list of SparseGraph tuples viewing the
    mappings (read-only arrays; a mapping lives as long as any array that views it)."""
    return [NpzGraphFile(f).graph for f in filenames]


# ---------------------------------------------------------------------------------------
# device batch
# ---------------------------------------------------------------------------------------
def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def _require_cuda(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.GnnsegError("gnn_fpga_b200 computes on CUDA devices only (got %s); there is no CPU path" % device)
    return device


class DeviceGraphBatch:
    """A batch of B graphs flattened into one block-diagonal graph on the GPU.

    n_nodes nodes in total; edge slot b*e_max + j is column j of event b in the padded
    (B, e_max) layout of merge_graphs (gnn/trainSegmentClassifier.py:66-95).  Slots that
    are padding carry src = dst = -1.
    """

    def __init__(self, X, src, dst, B, e_max, n_nodes_per_event=None, _prebuilt=None):
        dev = _require_cuda(X.device)
        assert X.dtype == torch.float32 and X.dim() == 2 and X.is_contiguous()
        assert src.dtype == torch.int32 and dst.dtype == torch.int32
        self.device = dev
        self.X = X
        self.src, self.dst = src, dst
        self.B, self.e_max = int(B), int(e_max)
        self.n_nodes = int(X.shape[0])
        self.n_slots = int(src.numel())
        assert self.n_slots == self.B * self.e_max and dst.numel() == self.n_slots
        self.F = int(X.shape[1])
        self.n_nodes_per_event = n_nodes_per_event
        self.n_real_edges = None       # filled lazily (needs a device->host read)
        self.node_perm = None          # set when the nodes were renumbered internally (GraphStore reorder)
        self._ws = {}
        self._graphs = {}
        if _prebuilt is None:
            self._build_csr()
            self.scores = torch.empty(self.n_slots, dtype=torch.float32, device=dev)
        else:                          # arrays assembled elsewhere (from_store): no CSR build here
            for k in ("in_ptr", "in_eid", "in_nbr", "in_pos", "out_ptr", "out_eid", "out_nbr", "out_pos", "scores", "adj_ptr", "adj",
                      "node_order", "status"):
                setattr(self, k, _prebuilt[k])
            self._ws = _prebuilt.get("ws", self._ws)
            self._set_struct()

    # -- construction ------------------------------------------------------------------
    def _build_csr(self):
        L = _lib.lib()
        dev, n, m = self.device, self.n_nodes, self.n_slots
        i32 = dict(dtype=torch.int32, device=dev)
        self.in_ptr = torch.empty(n + 1, **i32)
        self.out_ptr = torch.empty(n + 1, **i32)
        self.in_eid, self.in_nbr = torch.empty(m, **i32), torch.empty(m, **i32)
        self.out_eid, self.out_nbr = torch.empty(m, **i32), torch.empty(m, **i32)
        self.in_pos, self.out_pos = torch.empty(m, **i32), torch.empty(m, **i32)
        ws_bytes = 2 * L.gnnseg_csr_workspace_bytes(n, m)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_build_graph(_ptr(self.src), _ptr(self.dst), m, n,
                                            _ptr(self.in_ptr), _ptr(self.in_eid), _ptr(self.in_nbr), _ptr(self.in_pos),
                                            _ptr(self.out_ptr), _ptr(self.out_eid), _ptr(self.out_nbr), _ptr(self.out_pos),
                                            _ptr(ws), ws_bytes, _stream_ptr(dev)), "gnnseg_build_graph")
        self._csr_ws = ws   # keep alive until the stream has consumed it
        self._set_struct()

    def _set_struct(self):
        """GnnsegGraph over the arrays; builds the combined adjacency list of the fused inference path
        (gnnseg_build_adjacency: one launch on the current stream) if the arrays for it are there."""
        dev = self.device
        L = _lib.lib()
        built = getattr(self, "adj_ptr", None) is not None        # arrays assembled elsewhere (gnnseg_store_load_batch)
        if not built:
            self.adj_ptr = torch.empty(self.n_nodes + 1, dtype=torch.int32, device=dev)
            self.adj = torch.empty(L.gnnseg_adjacency_entries(self.n_nodes, self.n_slots), dtype=torch.int32, device=dev)
            self.node_order = torch.empty(max(self.n_nodes, 1), dtype=torch.int32, device=dev)
        if getattr(self, "status", None) is None:
            self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.struct = _lib.GnnsegGraph(
            self.n_nodes, self.n_slots, self.src.data_ptr(), self.dst.data_ptr(),
            self.in_ptr.data_ptr(), self.in_eid.data_ptr(), self.in_nbr.data_ptr(),
            self.out_ptr.data_ptr(), self.out_eid.data_ptr(), self.out_nbr.data_ptr(),
            self.in_pos.data_ptr(), self.out_pos.data_ptr(), self.adj_ptr.data_ptr(), self.adj.data_ptr(),
            self.node_order.data_ptr())
        if not built:
            with torch.cuda.device(dev):
                _lib.check(L.gnnseg_build_adjacency(C.byref(self.struct), _ptr(self.adj_ptr), _ptr(self.adj), _ptr(self.node_order),
                                                    _stream_ptr(dev)), "gnnseg_build_adjacency")

    @classmethod
    def from_dense(cls, X, Ri, Ro, validate=True):
        """X (B,N,F), Ri/Ro (B,N,E) fp32 CUDA tensors, as the reference's model receives them
        (gnn/model.py:142).  Raises ValueError for entries other than 0/1 or a column with
        more than one non-zero (the dense bmm would sum rows there; not a graph)."""
        dev = _require_cuda(X.device)
        if Ri.device != X.device or Ro.device != X.device:
            raise ValueError("X, Ri, Ro must live on the same device")
        if X.dim() != 3 or Ri.dim() != 3 or Ro.dim() != 3:
            raise ValueError("expected X (B,N,F), Ri (B,N,E), Ro (B,N,E)")
        B, N, F = X.shape
        E = Ri.shape[2]
        if tuple(Ri.shape) != (B, N, E) or tuple(Ro.shape) != (B, N, E):
            raise ValueError("Ri %s / Ro %s do not match X %s" % (tuple(Ri.shape), tuple(Ro.shape), tuple(X.shape)))
        X = X.to(torch.float32).contiguous()
        Ri = Ri.to(torch.float32).contiguous()
        Ro = Ro.to(torch.float32).contiguous()
        L = _lib.lib()
        src = torch.empty(B * E, dtype=torch.int32, device=dev)
        dst = torch.empty(B * E, dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gnnseg_dense_to_edges(_ptr(Ri), _ptr(Ro), B, N, E, _ptr(src), _ptr(dst),
                                               _ptr(err), _stream_ptr(dev)), "gnnseg_dense_to_edges")
        if validate:
            flag = int(err.item())
            if flag & _lib.BAD_VALUE:
                raise ValueError("Ri/Ro must contain only 0 and 1")
            if flag & _lib.BAD_HYPEREDGE:
                raise ValueError("a column of Ri or Ro has more than one non-zero entry")
        return cls(X.reshape(B * N, F), src, dst, B, E, n_nodes_per_event=[N] * B)

    @classmethod
    def from_graph_files(cls, filenames, device="cuda", pinned=None, n_threads=0):
        """A batch straight from graph files in the reference's NPZ format (gnn/graph.py:179-194):
        mapped, parsed and packed by the library, then as from_sparse_graphs."""
        dev = _require_cuda(torch.device(device))
        host = pack_npz_batch_host(list(filenames), pinned=pinned, n_threads=n_threads)
        return cls.from_packed_host(host, dev, pinned=pinned)

    @classmethod
    def from_sparse_graphs(cls, graphs, device="cuda", pinned=None, n_threads=0, reorder=False):
        """List of host SparseGraph tuples -> device batch, padded exactly as
        graph_from_sparse + merge_graphs would pad it (e_max = max len(Ri_rows); node rows
        are NOT padded because padded nodes influence no score).  reorder = True / "auto": through a
        GraphStore that renumbers the nodes for locality (gnn_fpga_b200/store.py)."""
        dev = _require_cuda(device)
        if reorder:
            from .store import GraphStore
            store = GraphStore.from_sparse_graphs(graphs, reorder=reorder, n_threads=n_threads)
            batch = cls.from_store(store, 0, len(store), dev)
            torch.cuda.current_stream(dev).synchronize()      # the temporary arena goes away with `store`
            return batch
        host = pack_sparse_batch_host(graphs, pinned=pinned, n_threads=n_threads)
        return cls.from_packed_host(host, dev, pinned=pinned)

    @classmethod
    def from_store(cls, store, lo=0, hi=None, device="cuda", bufs=None, copy_stream=None):
        """Events [lo, hi) of a GraphStore (gnn_fpga_b200/store.py): five contiguous asynchronous copies
        out of the store's pinned arena, then gnnseg_assemble_batch on the device (no sort, no atomics,
        no host work).  `bufs`: a DeviceBatchBuffers to reuse (pipelines); `copy_stream`: the stream the
        copies run on (the current stream waits for them before assembling)."""
        dev = _require_cuda(torch.device(device))
        hi = len(store) if hi is None else hi
        B = hi - lo
        if B <= 0:
            raise ValueError("empty batch")
        meta, n, e_max, n_in, n_out = store.batch_meta(lo, hi)
        if bufs is None:
            bufs = DeviceBatchBuffers(dev, n, n_in, n_out, B * e_max, B, store.F, store.col_bytes)
        v = bufs.views(n, n_in, n_out, B * e_max, B)
        compute = torch.cuda.current_stream(dev)
        cs = copy_stream if copy_stream is not None else compute
        with torch.cuda.device(dev):
            # six asynchronous copies out of the pinned arena on `cs`, then batch assembly and adjacency on the current stream
            _lib.check(_lib.lib().gnnseg_store_load_batch(C.byref(store.layout), store.arena.data_ptr(), lo, hi, C.byref(bufs.struct()),
                                                          C.c_void_p(cs.cuda_stream), C.c_void_p(compute.cuda_stream), None),
                       "gnnseg_store_load_batch")
        v["ws"] = bufs.ws
        batch = cls(v["X"], v["src"], v["dst"], B, e_max, n_nodes_per_event=np.diff(meta[:B + 1]).tolist(), _prebuilt=v)
        if store.order_column >= 0:
            batch.node_perm = (store, lo, hi)      # resolved lazily by original_node_ids(): only per-node outputs need it
        batch._bufs = bufs
        return batch

    def original_node_ids(self):
        """None, or an int64 device tensor `orig` with orig[i] = original flattened node id of internal node i
        (batches from a store that renumbered its nodes; per-edge scores never need it)."""
        if self.node_perm is None:
            return None
        if not isinstance(self.node_perm, torch.Tensor):
            store, lo, hi = self.node_perm
            n0, n1 = int(store.node_off[lo]), int(store.node_off[hi])
            local = store.perm[n0:n1].to(torch.int64)
            base = torch.from_numpy(np.repeat(store.node_off[lo:hi] - store.node_off[lo], np.diff(store.node_off[lo:hi + 1])))
            self.node_perm = (local + base).to(self.device)
        return self.node_perm

    @classmethod
    def from_packed_host(cls, host, device, pinned=None):
        """Second half of from_sparse_graphs: H2D of an already packed batch + device CSR build."""
        dev = _require_cuda(device)
        graphs = host["n_nodes"]
        X = host["X"].to(dev, non_blocking=True)
        src = host["src"].to(dev, non_blocking=True)
        dst = host["dst"].to(dev, non_blocking=True)
        if pinned is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            pinned["event"] = ev       # the caller must not refill the buffers before this fires
        return cls(X, src, dst, len(graphs), host["e_max"], n_nodes_per_event=host["n_nodes"])

    # -- helpers -------------------------------------------------------------------------
    def workspace(self, h):
        L = _lib.lib()
        nbytes = L.gnnseg_forward_workspace_bytes(self.n_nodes, self.n_slots, self.F, h)
        if nbytes == 0:
            _lib.check(-2, "gnnseg_forward_workspace_bytes(F=%d, h=%d)" % (self.F, h))
        if h not in self._ws or self._ws[h].numel() < nbytes:      # (shared pipeline buffers: grown to the largest batch)
            self._ws[h] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws[h]

    def train_workspace(self, h, n_iters):
        """Saved activations + backward scratch for gnnseg_forward_train / gnnseg_backward."""
        L = _lib.lib()
        key = ("train", h, n_iters)
        nbytes = L.gnnseg_train_workspace_bytes(self.n_nodes, self.n_slots, self.F, h, n_iters)
        if nbytes == 0:
            _lib.check(-2, "gnnseg_train_workspace_bytes(F=%d, h=%d, n_iters=%d)" % (self.F, h, n_iters))
        if key not in self._ws or self._ws[key].numel() < nbytes:
            self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws[key]

    def count_real_edges(self):
        if self.n_real_edges is None:
            self.n_real_edges = int(((self.src >= 0) & (self.dst >= 0)).sum().item())
        return self.n_real_edges

    def scores_2d(self):
        return self.scores.view(self.B, self.e_max)


class DeviceBatchBuffers:
    """Device buffers for batches up to a given size, reused from batch to batch (one per pipeline slot
    of predict_stream / batch_generator): input slices of the store, the assembled graph, the scores and
    the forward workspaces."""

    def __init__(self, device, n_nodes, n_in, n_out, n_slots, B, F, col_bytes):
        dev = _require_cuda(torch.device(device))
        i32 = dict(dtype=torch.int32, device=dev)
        self.cap = (n_nodes, n_in, n_out, n_slots, B)
        self.F, self.col_bytes = F, col_bytes
        col_dt = torch.int16 if col_bytes == 2 else torch.int32
        self.meta_host = torch.empty(3 * (B + 1), dtype=torch.int32, pin_memory=True)
        self.t = {
            "meta": torch.empty(3 * (B + 1), **i32),
            "X": torch.empty((max(n_nodes, 1), F), dtype=torch.float32, device=dev),
            "in_ptr_l": torch.empty(n_nodes + B, **i32), "out_ptr_l": torch.empty(n_nodes + B, **i32),
            "in_col": torch.empty(max(n_in, 1), dtype=col_dt, device=dev), "out_col": torch.empty(max(n_out, 1), dtype=col_dt, device=dev),
            "src": torch.empty(max(n_slots, 1), **i32), "dst": torch.empty(max(n_slots, 1), **i32),
            "in_pos": torch.empty(max(n_slots, 1), **i32), "out_pos": torch.empty(max(n_slots, 1), **i32),
            "in_ptr": torch.empty(n_nodes + 1, **i32), "out_ptr": torch.empty(n_nodes + 1, **i32),
            "in_eid": torch.empty(max(n_in, 1), **i32), "in_nbr": torch.empty(max(n_in, 1), **i32),
            "out_eid": torch.empty(max(n_out, 1), **i32), "out_nbr": torch.empty(max(n_out, 1), **i32),
            "scores": torch.empty(max(n_slots, 1), dtype=torch.float32, device=dev),
            "adj_ptr": torch.empty(n_nodes + 1, **i32),
            "adj": torch.empty(_lib.lib().gnnseg_adjacency_entries(n_nodes, n_slots), **i32),
            "node_order": torch.empty(max(n_nodes, 1), **i32),
            "status": torch.zeros(1, **i32),
        }
        self.ws = {}            # forward workspaces by hidden_dim, shared by the batches that pass through
        self.scores_host = None     # pinned result buffers (predict_stream): set by with_host_results()
        self.status_host = None
        self._struct = None

    def with_host_results(self):
        """Pinned host buffers the scores and the range flag of a batch come back to (gnnseg_store_forward_batch)."""
        if self.scores_host is None:
            self.scores_host = torch.empty(max(self.cap[3], 1), dtype=torch.float32, pin_memory=True)
            self.status_host = torch.zeros(1, dtype=torch.int32, pin_memory=True)
            self._struct = None
        return self

    def workspace(self, h):
        """Forward workspace for the largest batch these buffers hold."""
        if h not in self.ws:
            nbytes = _lib.lib().gnnseg_forward_workspace_bytes(self.cap[0], self.cap[3], self.F, h)
            if nbytes == 0:
                _lib.check(-2, "gnnseg_forward_workspace_bytes(F=%d, h=%d)" % (self.F, h))
            self.ws[h] = torch.empty(nbytes, dtype=torch.uint8, device=self.t["X"].device)
            self._struct = None
        return self.ws[h]

    def struct(self, h=None):
        """GnnsegBatchBuffers over the buffers (with the forward workspace for hidden_dim h when given)."""
        key = h
        if self._struct is None or self._struct[0] != key:
            t = self.t
            ws = self.workspace(h) if h is not None else None
            p = lambda x: x.data_ptr() if x is not None else None
            self._struct = (key, _lib.GnnsegBatchBuffers(
                self.cap[0], self.cap[1], self.cap[2], self.cap[3], self.cap[4], 0,
                p(t["meta"]), p(t["X"]), p(t["in_ptr_l"]), p(t["out_ptr_l"]), p(t["in_col"]), p(t["out_col"]), p(t["src"]), p(t["dst"]),
                p(t["in_pos"]), p(t["out_pos"]), p(t["in_ptr"]), p(t["out_ptr"]), p(t["adj_ptr"]), p(t["in_eid"]), p(t["in_nbr"]),
                p(t["out_eid"]), p(t["out_nbr"]), p(t["adj"]), p(t["node_order"]), p(t["scores"]), p(t["status"]), p(ws),
                ws.numel() if ws is not None else 0, p(self.meta_host), p(self.scores_host), p(self.status_host),
                getattr(self, "assemble_stream", None)))
        return self._struct[1]

    def fits(self, n_nodes, n_in, n_out, n_slots, B):
        return all(a <= b for a, b in zip((n_nodes, n_in, n_out, n_slots, B), self.cap))

    def views(self, n_nodes, n_in, n_out, n_slots, B):
        if not self.fits(n_nodes, n_in, n_out, n_slots, B):
            raise ValueError("batch %r exceeds the buffers %r" % ((n_nodes, n_in, n_out, n_slots, B), self.cap))
        t = self.t
        return {"meta": t["meta"][:3 * (B + 1)], "X": t["X"][:n_nodes], "in_ptr_l": t["in_ptr_l"][:n_nodes + B],
                "out_ptr_l": t["out_ptr_l"][:n_nodes + B], "in_col": t["in_col"][:n_in], "out_col": t["out_col"][:n_out],
                "src": t["src"][:n_slots], "dst": t["dst"][:n_slots], "in_pos": t["in_pos"][:n_slots], "out_pos": t["out_pos"][:n_slots],
                "in_ptr": t["in_ptr"][:n_nodes + 1], "out_ptr": t["out_ptr"][:n_nodes + 1], "in_eid": t["in_eid"][:n_in],
                "in_nbr": t["in_nbr"][:n_in], "out_eid": t["out_eid"][:n_out], "out_nbr": t["out_nbr"][:n_out],
                "scores": t["scores"][:n_slots], "adj_ptr": t["adj_ptr"][:n_nodes + 1], "adj": t["adj"],
                "node_order": t["node_order"][:max(n_nodes, 1)], "status": t["status"]}


def pack_npz_batch_host(filenames, pinned=None, n_threads=0):
    """Graph files (the reference's NPZ format) -> packed host batch, without building a Python
    object per array: the library maps and parses the files (gnnseg_npz_open_batch_host), the packer
    reads the mappings (gnnseg_pack_sparse_batch_host), the files are unmapped again.  Returns what
    pack_sparse_batch_host returns."""
    import ctypes as C
    if n_threads <= 0:
        import os
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        n_threads = max(1, (os.cpu_count() or 4) // local_world - 2)
    L = _lib.lib()
    B = len(filenames)
    if B == 0:
        raise ValueError("empty batch")
    paths = (C.c_char_p * B)(*[str(f).encode() for f in filenames])
    gs = (_lib.GnnsegNpzGraph * B)()
    rc = L.gnnseg_npz_open_batch_host(B, C.addressof(paths), C.addressof(gs), n_threads)
    _lib.check(rc, "gnnseg_npz_open_batch_host")
    try:
        rec = np.frombuffer(gs, dtype=np.dtype([(n, "<u8" if t is C.c_void_p else ("<i8" if t is C.c_int64 else "<i4"))
                                                for n, t in _lib.GnnsegNpzGraph._fields_], align=True))
        F = int(rec["n_features"][0])
        if not np.all(rec["n_features"] == F):
            raise ValueError("graph files with different numbers of features in one batch")
        n_nodes = np.ascontiguousarray(rec["n_nodes"])
        n_in, n_out = np.ascontiguousarray(rec["n_in"]), np.ascontiguousarray(rec["n_out"])
        e_max, nt = int(n_in.max()), int(n_nodes.sum())
        can_pin = torch.cuda.is_available()
        if pinned is not None:
            # a reusable staging dict (see SegmentClassifier._grow_pinned): grown in place when too small
            if pinned.get("X") is None or pinned["X"].shape[0] < nt or pinned["X"].shape[1] != F or pinned["src"].numel() < B * e_max:
                cap_n, cap_m = int(nt * 1.25) + 1, int(B * e_max * 1.25) + 1
                pinned["X"] = torch.empty((cap_n, F), dtype=torch.float32, pin_memory=can_pin)
                pinned["src"] = torch.empty(cap_m, dtype=torch.int32, pin_memory=can_pin)
                pinned["dst"] = torch.empty(cap_m, dtype=torch.int32, pin_memory=can_pin)
            Xo, src, dst = pinned["X"][:nt], pinned["src"][:B * e_max], pinned["dst"][:B * e_max]
        else:
            Xo = torch.empty((nt, F), dtype=torch.float32, pin_memory=can_pin)
            src = torch.empty(B * e_max, dtype=torch.int32, pin_memory=can_pin)
            dst = torch.empty(B * e_max, dtype=torch.int32, pin_memory=can_pin)
        cols = [np.ascontiguousarray(rec[k]) for k in ("X", "Ri_rows", "Ri_cols", "Ro_rows", "Ro_cols")]
        ptr = lambda a: a.__array_interface__["data"][0]
        rc = L.gnnseg_pack_sparse_batch_host(B, F, e_max, ptr(cols[0]), ptr(n_nodes), ptr(cols[1]), ptr(cols[2]), ptr(cols[3]),
                                             ptr(cols[4]), ptr(n_in), ptr(n_out), Xo.data_ptr(), src.data_ptr(), dst.data_ptr(),
                                             n_threads)
    finally:
        L.gnnseg_npz_close_batch_host(B, C.addressof(gs))
    if rc == -1:
        raise ValueError("graph file index out of range (row >= n_nodes or col >= max len(Ri_rows))")
    if rc == _lib.EHYPEREDGE:
        raise ValueError("a column of Ri or Ro has more than one non-zero entry")
    _lib.check(rc, "gnnseg_pack_sparse_batch_host")
    return {"X": Xo, "src": src, "dst": dst, "e_max": e_max, "n_nodes": n_nodes.tolist()}


def pack_sparse_batch_host(graphs, pinned=None, n_threads=0):
    """Run gnnseg_pack_sparse_batch_host over a list of SparseGraph tuples.  Returns pinned
    host tensors X (n_nodes,F) f32, src/dst (B*e_max) i32.  `pinned` may pass a dict of
    pre-allocated (larger) pinned buffers to reuse.  n_threads <= 0: this process's share of
    the host cores (cores / LOCAL_WORLD_SIZE) minus two."""
    if n_threads <= 0:
        import os
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        n_threads = max(1, (os.cpu_count() or 4) // local_world - 2)
    L = _lib.lib()
    B = len(graphs)
    if B == 0:
        raise ValueError("empty batch")
    F = int(graphs[0].X.shape[1])
    keep = []   # keep converted arrays alive during the call

    def arr(a, dt):
        if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
        return a

    Xs = [arr(g.X, np.float32) for g in graphs]
    cols = [[arr(g[k], np.int64) for g in graphs] for k in (1, 2, 3, 4)]    # Ri_rows, Ri_cols, Ro_rows, Ro_cols
    n_nodes = np.fromiter((x.shape[0] for x in Xs), dtype=np.int64, count=B)
    n_in = np.fromiter((a.shape[0] for a in cols[0]), dtype=np.int64, count=B)
    n_out = np.fromiter((a.shape[0] for a in cols[2]), dtype=np.int64, count=B)
    for b in range(B):
        if Xs[b].ndim != 2 or Xs[b].shape[1] != F:
            raise ValueError("graph %d: X must be (N, %d)" % (b, F))
        if cols[1][b].shape[0] != n_in[b] or cols[3][b].shape[0] != n_out[b]:
            raise ValueError("graph %d: rows/cols length mismatch" % b)
    e_max = int(n_in.max())    # graph_from_sparse: n_edges = len(Ri_rows)
    nt = int(n_nodes.sum())
    can_pin = torch.cuda.is_available()
    if pinned is not None:
        Xo, src, dst = pinned["X"][:nt], pinned["src"][:B * e_max], pinned["dst"][:B * e_max]
    else:
        Xo = torch.empty((nt, F), dtype=torch.float32, pin_memory=can_pin)
        src = torch.empty(B * e_max, dtype=torch.int32, pin_memory=can_pin)
        dst = torch.empty(B * e_max, dtype=torch.int32, pin_memory=can_pin)

    def pp(arrs):     # array of B data pointers (numpy's array interface is much cheaper than .ctypes)
        t = np.fromiter((a.__array_interface__["data"][0] for a in arrs), dtype=np.uintp, count=B)
        keep.append(t)
        return t.__array_interface__["data"][0]

    rc = L.gnnseg_pack_sparse_batch_host(
        B, F, e_max, pp(Xs), n_nodes.__array_interface__["data"][0], pp(cols[0]), pp(cols[1]), pp(cols[2]), pp(cols[3]),
        n_in.__array_interface__["data"][0], n_out.__array_interface__["data"][0],
        Xo.data_ptr(), src.data_ptr(), dst.data_ptr(), n_threads)
    if rc == -1:
        raise ValueError("SparseGraph index out of range (row >= n_nodes or col >= max len(Ri_rows))")
    if rc == _lib.EHYPEREDGE:
        raise ValueError("a column of Ri or Ro has more than one non-zero entry")
    _lib.check(rc, "gnnseg_pack_sparse_batch_host")
    return {"X": Xo, "src": src, "dst": dst, "e_max": e_max, "n_nodes": n_nodes.tolist()}
