"""gnn_fpga_b200: B200-native implementation of gnn-fpga's SegmentClassifier hot path.

Public surface mirrors the reference modules gnn/model.py, gnn/graph.py:
    SegmentClassifier, EdgeNetwork, NodeNetwork, MaskedLinear,
    Graph, SparseGraph, make_sparse_graph, graph_from_sparse, save_graph(s), load_graph(s)
plus the device batch (DeviceGraphBatch) that replaces the dense incidence tensors, the load-time
event store (GraphStore) and the reference's generator protocol on top of it (batch_generator).
"""
from .graph import (Graph, SparseGraph, make_sparse_graph, graph_from_sparse, save_graph,  # noqa: F401
                    save_graphs, load_graph, load_graphs, load_graphs_mapped, NpzGraphFile, DeviceGraphBatch,
                    pack_sparse_batch_host, pack_npz_batch_host)
from .store import GraphStore, StoreBatch  # noqa: F401
from .loader import batch_generator  # noqa: F401
from .model import MaskedLinear, EdgeNetwork, NodeNetwork, SegmentClassifier  # noqa: F401
from ._lib import GnnsegError, LIB_PATH  # noqa: F401
