"""The C-ABI library loads and exports every symbol include/gnnseg.h declares.  No compute
calls here (no GPU needed): only the pure-host queries are invoked."""
import ctypes
import os
import re

from conftest import ROOT
from gnn_fpga_b200 import _lib


def declared_functions():
    text = open(os.path.join(ROOT, "include", "gnnseg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnnseg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    names = declared_functions()
    assert "gnnseg_forward" in names and "gnnseg_build_csr" in names and len(names) >= 14


def test_every_declared_symbol_is_exported_and_bound():
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    for n in names:
        assert hasattr(handle, n), "libgnnseg_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"


def test_host_only_queries():
    L = _lib.lib()
    assert L.gnnseg_abi_version() == _lib.ABI_VERSION == 5
    assert L.gnnseg_strerror(0) == b"ok"
    assert b"unsupported" in L.gnnseg_strerror(-2)
    for h in (4, 8, 16, 32, 64):
        for F in (1, 2, 3, 4):
            assert L.gnnseg_supported(F, h) == 1
    for F, h in ((0, 8), (5, 8), (3, 12), (3, 128), (3, 0)):
        assert L.gnnseg_supported(F, h) == 0
        assert L.gnnseg_weights_floats(F, h) == 0
    # blob size: Win^T[4][h] + b_in + WP^T[(h+4)][5h] + bias[5h] + W2 + 4 + W4^T + b4
    #            + tf32 hi/lo operand images of W4 [h][h] and of WP [5h][40]
    #            + the fused path's edge constants (-2 w2, b2 + sum w2, 2^(log2e b1))
    h = 32
    assert L.gnnseg_weights_floats(3, h) == (4 * h + h + (h + 4) * 5 * h + 5 * h + h + 4 + h * h + h
                                             + 2 * h * h + 2 * 5 * h * 40 + 2 * h + 4)
    # workspace: X4 + P + max(2 Q + 2 e, 2 x (n + 1) state rows of 5h + status)
    assert L.gnnseg_forward_workspace_bytes(1000, 5000, 3, 32) >= 4 * (1000 * 4 + 1000 * 64 + 2 * 1001 * 160)
    assert L.gnnseg_adjacency_entries(1000, 5000) >= 2 * 5000 + 4 * 1000
    assert L.gnnseg_forward_workspace_bytes(10, 5000, 3, 8) >= 4 * (10 * 4 + 10 * 16 + 2 * 10 * 24 + 2 * 5000)
    assert L.gnnseg_forward_workspace_bytes(1000, 5000, 3, 12) == 0
    assert L.gnnseg_csr_workspace_bytes(1000, 5000) >= 8 * 1001


def test_argument_errors_without_device_work():
    """Null / unsupported arguments are rejected on the host before any CUDA call."""
    L = _lib.lib()
    assert L.gnnseg_pack_weights(None, 3, 32, None, None) == -1
    assert L.gnnseg_pack_weights(None, 3, 12, None, None) == -2
    assert L.gnnseg_forward(None, None, None, 3, 12, 1, None, None, 0, None) == -2
    assert L.gnnseg_forward(None, None, None, 3, 32, 1, None, None, 0, None) == -1
    assert L.gnnseg_dense_to_edges(None, None, -1, 1, 1, None, None, None, None) == -1
    assert L.gnnseg_build_csr(None, None, 5, 5, None, None, None, None, None, 0, None) == -1
    # the fused inference path and the event store
    assert L.gnnseg_forward_ex(None, None, None, 3, 32, 1, None, None, 0, 0, None, None) == -1
    assert L.gnnseg_build_adjacency(None, None, None, None, None) == -1
    assert L.gnnseg_store_batch_shape_host(None, None, 0, 1, None) == -1
    assert L.gnnseg_store_load_batch(None, None, 0, 1, None, None, None, None) == -1
    assert L.gnnseg_store_forward_batch(None, None, 0, 1, None, 32, 1, 0, None, None, None, None, None) == -1
    assert L.gnnseg_fused_gather_step(None, None, None, 16, None, 16, None) == -2       # hidden_dim 32 / 64 only
    assert L.gnnseg_fused_gather_step(None, None, None, 32, None, 32, None) == -1
    assert L.gnnseg_edge_final_step(None, None, None, 64, 0, 32, 32, None, None) == -1
    assert L.gnnseg_state_input_step(None, None, 10, 3, 32, None, None, 100, None, None) == -1    # n_cols: 5h or 2h
    assert L.gnnseg_state_mlp_step(None, None, None, 32, 10, 8, None, 40, None, None) == -2
    assert L.gnnseg_assemble_batch(None, 2, 10, 5, 3, 3, None, None, None, None, 3, None, None, None, None, None, None,
                                   None, None, None, None, None) == -1                             # col_bytes 2 or 4
    assert L.gnnseg_store_plan_host(1, 0, None, None, None, None, None, None) == -1
    assert b"hyper-edge" in L.gnnseg_strerror(-8)


def test_training_and_segment_entry_points_check_arguments_on_the_host():
    """The entry points added for the training step and for segment construction reject bad
    arguments before any device work; their size queries are pure host functions."""
    L = _lib.lib()
    n, m, F, h, T = 1000, 5000, 3, 32, 4
    need = 4 * (n * 4 + (T + 1) * n * h + T * n * h + (T + 1) * n * 2 * h + T * n * 3 * h + 2 * T * m      # saved by the forward
                + n * h + n * 5 * h + 2 * m)                                                                # backward scratch
    assert L.gnnseg_train_workspace_bytes(n, m, F, h, T) >= need
    assert L.gnnseg_train_workspace_bytes(n, m, F, 12, T) == 0 and L.gnnseg_train_workspace_bytes(n, m, F, h, 65) == 0
    assert L.gnnseg_forward_train(None, None, None, F, h, T, None, None, 0, None) == -1
    assert L.gnnseg_forward_train(None, None, None, F, 12, T, None, None, 0, None) == -2
    assert L.gnnseg_backward(None, None, None, F, h, T, None, None, None, 0, None) == -1
    assert L.gnnseg_backward(None, None, None, F, 5, T, None, None, None, 0, None) == -2
    assert L.gnnseg_bce_loss(None, None, None, 10, None, None, None, None) == -1
    assert L.gnnseg_l1_penalty(None, F, h, 0.1, None, None, None) == -1
    assert L.gnnseg_adam_step(None, None, None, None, 10, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, None) == -1     # step counts from 1
    assert L.gnnseg_node_gather_step(None, None, None, None, h, None, h, None) == -1
    assert L.gnnseg_node_mlp_step(None, None, None, h, 10, h, None, None, None) == -1
    assert L.gnnseg_node_mlp_step(None, None, None, 12, 10, 12, None, None, None) == -2
    assert L.gnnseg_segments_workspace_bytes(4000, 9) >= 4 * (4000 + 9 * 4000)
    assert L.gnnseg_segments_workspace_bytes(4000, 33) == 0
    assert L.gnnseg_build_segments(None, None, None, None, 2, None, 0, None, 0, 1, 1e-3, 1e-3, 200.0, 5, 0, 0,
                                   None, None, None, None, None, 0, None) == -2                            # dtype
    assert L.gnnseg_build_segments(None, None, None, None, 4, None, 0, None, 0, 1, 1e-3, 1e-3, 200.0, 5, 0, 0,
                                   None, None, None, None, None, 0, None) == -1                            # no n_edges word
    assert L.gnnseg_scale_features(None, None, None, 4, 10, 1.0, 1.0, 1.0, None, None) == -1
    assert b"npz" in L.gnnseg_strerror(-7) and b"mapped" in L.gnnseg_strerror(-6)


def test_node_classifier_entry_points_check_arguments_on_the_host():
    L = _lib.lib()
    F, h, T = 4, 32, 2
    assert L.gnnseg_pack_node_head(None, None, None, F, h, None, None) == -1
    assert L.gnnseg_pack_node_head(None, None, None, 5, h, None, None) == -2
    assert L.gnnseg_forward_nodes(None, None, None, None, F, h, T, None, None, 0, None) == -1
    assert L.gnnseg_forward_nodes(None, None, None, None, F, 12, T, None, None, 0, None) == -2
    assert L.gnnseg_forward_nodes_train(None, None, None, None, F, h, 65, None, None, 0, None) == -1
    assert L.gnnseg_backward_nodes(None, None, None, None, F, h, T, None, None, None, None, None, 0, None) == -1
    assert L.gnnseg_backward_nodes(None, None, None, None, F, 7, T, None, None, None, None, None, 0, None) == -2


def test_batched_segment_entry_point_checks_arguments_on_the_host():
    L = _lib.lib()
    assert L.gnnseg_segments_batch_workspace_bytes(64, 256000, 9) >= 4 * (64 * 68 + 256000 + 9 * 256000 + 64)
    assert L.gnnseg_segments_batch_workspace_bytes(1, 1000, 33) == 0
    args = lambda **k: [None, None, None, None, k.get("nb", 4), None, k.get("B", 2), None, 100, 60, None, 0, k.get("nl", 10),
                        1.0, 1.0, 1.0, 5, 0, None, None, None, None, None, 0, None]
    assert L.gnnseg_build_segments_batch(*args()) == -1            # no workspace / columns
    assert L.gnnseg_build_segments_batch(*args(nb=2)) == -2        # dtype
    assert L.gnnseg_build_segments_batch(*args(nl=40)) == -1
