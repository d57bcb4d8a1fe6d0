"""ctypes binding of libgnnseg_b200.so (the C ABI in include/gnnseg.h).

There is no CPU fallback: if the shared library is missing, `lib()` raises, and every
compute entry point of the package goes through `lib()`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GNNSEG_LIB: another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.path.abspath(os.environ["GNNSEG_LIB"]) if os.environ.get("GNNSEG_LIB") else os.path.join(_HERE, "libgnnseg_b200.so")

OK = 0
ABI_VERSION = 5
ERRORS = {-1: "EINVAL", -2: "EUNSUPPORTED", -3: "EWORKSPACE", -4: "ECUDA", -5: "ENODEVICE", -6: "EIO", -7: "EFORMAT", -8: "EHYPEREDGE"}
EHYPEREDGE = -8
BAD_VALUE = 1
BAD_HYPEREDGE = 2
FWD_EXACT = 1            # gnnseg_forward_ex flags: force the step-by-step kernels
RANGE_CLAMPED = 1        # gnnseg_forward_ex status bit: an edge projection left the fused path's range

_f32p = C.c_void_p
_i32p = C.c_void_p


class GnnsegParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "w_in", "b_in", "w_e1", "b_e1", "w_e2", "b_e2", "w_n1", "b_n1", "w_n2", "b_n2",
        "m_e1", "m_e2", "m_n1", "m_n2")]


class GnnsegGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "w_in", "b_in", "w_e1", "b_e1", "w_e2", "b_e2", "w_n1", "b_n1", "w_n2", "b_n2")]


class GnnsegNpzGraph(C.Structure):
    _fields_ = [("handle", C.c_void_p), ("X", C.c_void_p), ("n_nodes", C.c_int64), ("n_features", C.c_int32),
                ("Ri_rows", C.c_void_p), ("Ri_cols", C.c_void_p), ("n_in", C.c_int64),
                ("Ro_rows", C.c_void_p), ("Ro_cols", C.c_void_p), ("n_out", C.c_int64),
                ("y", C.c_void_p), ("n_y", C.c_int64)]


class GnnsegStoreLayout(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("n_features", C.c_int32), ("col_bytes", C.c_int32)] + [(n, C.c_int64) for n in (
        "total_nodes", "total_in", "total_out", "total_y", "o_node_off", "o_in_off", "o_out_off", "o_y_off",
        "o_X", "o_in_ptr", "o_out_ptr", "o_in_col", "o_out_col", "o_y", "o_perm", "o_n_edges", "bytes")]


class GnnsegBatchBuffers(C.Structure):
    """One pipeline slot of gnnseg_store_forward_batch (include/gnnseg.h)."""
    _fields_ = ([(n, C.c_int32) for n in ("cap_nodes", "cap_in", "cap_out", "cap_slots", "cap_events", "reserved_")] +
                [(n, C.c_void_p) for n in ("meta", "X", "in_ptr_local", "out_ptr_local", "in_col", "out_col", "src", "dst", "in_pos",
                                           "out_pos", "in_ptr", "out_ptr", "adj_ptr", "in_eid", "in_nbr", "out_eid", "out_nbr", "adj",
                                           "node_order", "scores", "status", "ws")] +
                [("ws_bytes", C.c_size_t)] + [(n, C.c_void_p) for n in ("meta_host", "scores_host", "status_host", "assemble_stream")])


class GnnsegGraph(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_slots", C.c_int32)] + [(n, C.c_void_p) for n in (
        "src", "dst", "in_ptr", "in_eid", "in_nbr", "out_ptr", "out_eid", "out_nbr", "in_pos", "out_pos", "adj_ptr", "adj", "node_order")]


EWORKSPACE = -3

# name -> (restype, argtypes); must list every symbol include/gnnseg.h declares
SIGNATURES = {
    "gnnseg_abi_version": (C.c_int, []),
    "gnnseg_strerror": (C.c_char_p, [C.c_int]),
    "gnnseg_supported": (C.c_int, [C.c_int, C.c_int]),
    "gnnseg_device_sm_count": (C.c_int, []),
    "gnnseg_weights_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "gnnseg_pack_weights": (C.c_int, [C.POINTER(GnnsegParams), C.c_int, C.c_int, _f32p, C.c_void_p]),
    "gnnseg_dense_to_edges": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, C.c_void_p]),
    "gnnseg_csr_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gnnseg_build_csr": (C.c_int, [_i32p, _i32p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_build_graph": (C.c_int, [_i32p, _i32p, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_forward_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gnnseg_forward": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_forward_ex": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t,
                                   C.c_int, _i32p, C.c_void_p]),
    "gnnseg_adjacency_entries": (C.c_size_t, [C.c_int, C.c_int]),
    "gnnseg_build_adjacency": (C.c_int, [C.POINTER(GnnsegGraph), _i32p, _i32p, _i32p, C.c_void_p]),
    "gnnseg_store_batch_shape_host": (C.c_int, [C.POINTER(GnnsegStoreLayout), C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "gnnseg_store_load_batch": (C.c_int, [C.POINTER(GnnsegStoreLayout), C.c_void_p, C.c_int, C.c_int, C.POINTER(GnnsegBatchBuffers),
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnnseg_store_forward_batch": (C.c_int, [C.POINTER(GnnsegStoreLayout), C.c_void_p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(GnnsegBatchBuffers), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gnnseg_state_input_step": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _i32p, C.c_void_p]),
    "gnnseg_fused_gather_step": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, _f32p, C.c_int, C.c_void_p]),
    "gnnseg_state_mlp_step": (C.c_int, [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, _i32p, C.c_void_p]),
    "gnnseg_edge_final_step": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p]),
    "gnnseg_input_step": (C.c_int, [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_void_p]),
    "gnnseg_edge_step": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, _f32p, _f32p, _f32p, C.c_void_p]),
    "gnnseg_node_step": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, _f32p, _f32p, _f32p, C.c_int, _f32p, _f32p, C.c_void_p]),
    "gnnseg_node_gather_step": (C.c_int, [C.POINTER(GnnsegGraph), _f32p, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.c_void_p]),
    "gnnseg_node_mlp_step": (C.c_int, [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_void_p]),
    "gnnseg_train_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gnnseg_forward_train": (C.c_int, [_f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_backward": (C.c_int, [_f32p, C.POINTER(GnnsegParams), C.POINTER(GnnsegGraph), C.c_int, C.c_int, C.c_int, _f32p,
                                 C.POINTER(GnnsegGrads), C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_pack_node_head": (C.c_int, [C.POINTER(GnnsegParams), _f32p, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p]),
    "gnnseg_forward_nodes": (C.c_int, [_f32p, _f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p,
                                      C.c_size_t, C.c_void_p]),
    "gnnseg_forward_nodes_train": (C.c_int, [_f32p, _f32p, C.POINTER(GnnsegGraph), _f32p, C.c_int, C.c_int, C.c_int, _f32p,
                                            C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_backward_nodes": (C.c_int, [_f32p, _f32p, C.POINTER(GnnsegParams), C.POINTER(GnnsegGraph), C.c_int, C.c_int, C.c_int,
                                       _f32p, C.POINTER(GnnsegGrads), _f32p, _f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_bce_loss": (C.c_int, [_f32p, _f32p, _f32p, C.c_int, _f32p, _f32p, C.c_void_p, C.c_void_p]),
    "gnnseg_l1_penalty": (C.c_int, [C.POINTER(GnnsegParams), C.c_int, C.c_int, C.c_float, _f32p, C.POINTER(GnnsegGrads), C.c_void_p]),
    "gnnseg_adam_step": (C.c_int, [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_void_p]),
    "gnnseg_segments_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gnnseg_segments_batch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gnnseg_build_segments_batch": (C.c_int, [_i32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, _i32p, C.c_int,
                                             C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                             C.c_int, _i32p, _i32p, _f32p, _i32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_build_segments": (C.c_int, [_i32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f32p,
                                       _i32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gnnseg_scale_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                       _f32p, C.c_void_p]),
    "gnnseg_npz_open_graph_host": (C.c_int, [C.c_char_p, C.POINTER(GnnsegNpzGraph)]),
    "gnnseg_npz_close_graph_host": (C.c_int, [C.POINTER(GnnsegNpzGraph)]),
    "gnnseg_npz_open_batch_host": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "gnnseg_npz_close_batch_host": (C.c_int, [C.c_int, C.c_void_p]),
    "gnnseg_store_plan_host": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(GnnsegStoreLayout)]),
    "gnnseg_store_fill_host": (C.c_int, [C.POINTER(GnnsegStoreLayout), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "gnnseg_assemble_batch": (C.c_int, [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, C.c_void_p, C.c_void_p,
                                       C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_void_p]),
    "gnnseg_pack_sparse_batch_host": (C.c_int, [
        C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
}

_lib = None


class GnnsegError(RuntimeError):
    pass


def lib():
    """Load the shared library once.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GnnsegError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gnn_fpga_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.gnnseg_abi_version() != ABI_VERSION:
            raise GnnsegError("ABI version mismatch in %s" % LIB_PATH)
        _lib = handle
    return _lib


def check(rc, what):
    if rc != OK:
        msg = lib().gnnseg_strerror(rc).decode()
        raise GnnsegError("%s failed: %s (%s)" % (what, msg, ERRORS.get(rc, rc)))
