"""The oracle against outputs of the reference itself (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import MODEL_CASES, NODECLF_CASES, load_case, rel_err
from oracle import segclf_oracle as O


@pytest.mark.parametrize("name", MODEL_CASES)
def test_dense_restatement_matches_reference(name):
    rec = load_case(name)
    torch.set_num_threads(1)
    p = O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"])
    out = O.dense_forward(p, torch.from_numpy(rec["X"]), torch.from_numpy(rec["Ri"].astype(np.float32)),
                          torch.from_numpy(rec["Ro"].astype(np.float32)), rec["n_iters"])
    assert out.shape == rec["out"].shape
    # same torch ops in the same order: equal to the last bit in practice; 1e-6 leaves room for
    # a different BLAS blocking on another host
    assert rel_err(out.numpy(), rec["out"]) <= 1e-6


@pytest.mark.parametrize("name", MODEL_CASES)
def test_sparse_restatement_matches_reference(name):
    rec = load_case(name)
    p = O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"])
    B, N, F = rec["X"].shape
    src, dst = O.edges_from_dense(rec["Ri"], rec["Ro"])
    out32 = O.sparse_forward(p, rec["X"].reshape(B * N, F), src, dst, rec["n_iters"], torch.float32)
    out64 = O.sparse_forward(p, rec["X"].reshape(B * N, F), src, dst, rec["n_iters"], torch.float64)
    ref = rec["out"].reshape(-1)
    assert rel_err(out32.numpy(), ref) <= 2e-6
    assert rel_err(out64.numpy(), ref) <= 2e-6


def test_init_params_reproduces_reference_init():
    """init_params(F, h, seed) draws the same numbers as the reference constructor."""
    rec = load_case("c1_toy2d_h8_it1")
    p = O.init_params(rec["F"], rec["h"], rec["seed"])
    for k in O.PARAM_KEYS:
        assert torch.equal(p[k], rec["params"][k]), k


def test_param_counts_and_keys():
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "structure.npz"))
    assert tuple(z["keys"]) == O.PARAM_KEYS
    for key, count in z["counts"]:
        F, h = map(int, key.split(","))
        assert sum(v.numel() for v in O.init_params(F, h).values()) == int(count)
    # 189 / 569 / 6881 / 26049 / 6689
    assert dict((k, int(v)) for k, v in z["counts"]) == {"3,4": 189, "3,8": 569, "3,32": 6881, "3,64": 26049, "2,32": 6689}


def test_integer_oracle_matches_reference_nonzero():
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "graph_roundtrip.npz"))
    src, dst = O.edges_from_dense(z["Ri"][None], z["Ro"][None])
    n = z["Ri"].shape[0]
    ptr, eid = O.csr_from_keys(dst, n)
    assert np.array_equal(eid, z["Ri_cols"])
    assert np.array_equal(np.repeat(np.arange(n), np.diff(ptr)), z["Ri_rows"])
    ptr, eid = O.csr_from_keys(src, n)
    assert np.array_equal(eid, z["Ro_cols"])
    assert np.array_equal(np.repeat(np.arange(n), np.diff(ptr)), z["Ro_rows"])


def test_padding_constant():
    """A -1/-1 slot scores sigmoid(W2.tanh(b1)+b2) at every edge step."""
    rec = load_case("acts_ragged_h32_it4")
    p = rec["params"]
    const = torch.sigmoid(p[O.PARAM_KEYS[4]] @ torch.tanh(p[O.PARAM_KEYS[3]]) + p[O.PARAM_KEYS[5]]).item()
    pad = (rec["Ri"].sum(axis=1) == 0) & (rec["Ro"].sum(axis=1) == 0)
    assert pad.any()
    assert np.allclose(rec["out"][pad], const, rtol=1e-6)


@pytest.mark.parametrize("name", ["train_step_h8_it2", "train_grads_h32_it3"])
def test_gradient_oracle_matches_reference_estimator(name):
    """sparse_vjp (autograd over the sparse restatement, the checker of gnnseg_backward) against
    the loss and first-step gradients recorded from the reference's own Estimator.training_step
    (tests/golden/train_step_h8_it2.npz, oracle/make_golden.py; train_grads_h32_it3.npz,
    oracle/make_golden_train32.py: the width whose CUDA backward runs on the tensor cores)."""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = {k[6:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("param:")}
    src, dst = O.edges_from_dense(z["Ri"], z["Ro"])
    B, N, F = z["X"].shape
    _, loss, g = O.sparse_vjp(p, z["X"].reshape(B * N, F), src, dst, int(z["n_iters"]), y=z["y"], l1=float(z["l1"]))
    assert abs(float(loss) - float(z["losses"][0])) <= 1e-6 * abs(float(z["losses"][0]))
    for k, v in g.items():
        ref = z["grad0:" + k]
        assert np.max(np.abs(v.numpy() - ref)) <= 2e-5 * np.max(np.abs(ref)), k


SEGMENT_CASES = ["segments_f32_small", "segments_f32_event", "segments_f64_small", "segments_f32_wide"]


@pytest.mark.parametrize("name", SEGMENT_CASES)
def test_segment_oracle_matches_reference_construct_graph(name):
    """oracle/segments_oracle.py (numpy, no pandas) against the edges, labels and features the
    reference's own construct_graph produced (oracle/make_golden_segments.py): bit-exact."""
    import os
    from conftest import GOLDEN
    from oracle import segments_oracle as S
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    s, e, y = S.build_segments(z["layer"], z["r"], z["phi"], z["z"], z["particle_id"], z["layer_pairs"],
                               float(z["phi_slope_max"]), float(z["phi_slope_outer_max"]), float(z["z0_max"]))
    assert np.array_equal(s, z["seg_start"]) and np.array_equal(e, z["seg_end"]) and np.array_equal(y, z["y"])
    assert np.array_equal(S.features([z["r"], z["phi"], z["z"]], z["feature_scale"]), z["X"])


# ---- NodeClassifier (gnn/MPNN_HitClassifier.ipynb c21; goldens from oracle/make_golden_nodeclf.py) ----
@pytest.mark.parametrize("name", NODECLF_CASES)
def test_nodeclf_restatements_match_reference(name):
    rec = load_case(name)
    torch.set_num_threads(1)
    p = rec["params"]
    assert tuple(rec["keys"]) == O.PARAM_KEYS + O.HEAD_KEYS
    X = torch.from_numpy(rec["X"])
    out = O.nodeclf_dense_forward(p, X, torch.from_numpy(rec["Ri"].astype(np.float32)),
                                  torch.from_numpy(rec["Ro"].astype(np.float32)), rec["n_iters"])
    assert out.shape == rec["out"].shape
    assert rel_err(out.numpy(), rec["out"]) <= 1e-6
    B, N, F = rec["X"].shape
    src, dst = O.edges_from_dense(rec["Ri"], rec["Ro"])
    out32 = O.nodeclf_sparse_forward(p, rec["X"].reshape(B * N, F), src, dst, rec["n_iters"])
    assert rel_err(out32.numpy(), rec["out"].reshape(-1)) <= 2e-6


@pytest.mark.parametrize("name", NODECLF_CASES)
def test_nodeclf_gradient_oracle_matches_reference(name):
    """BCELoss backward through the notebook's own model (recorded) vs autograd over the restatement."""
    rec = load_case(name)
    B, N, F = rec["X"].shape
    src, dst = O.edges_from_dense(rec["Ri"], rec["Ro"])
    _, loss, grads = O.nodeclf_sparse_vjp(rec["params"], rec["X"].reshape(B * N, F), src, dst, rec["n_iters"], y=rec["y"])
    assert abs(float(loss) - float(rec["loss"])) <= 1e-6 * abs(float(rec["loss"]))
    for k, ref in rec["grads"].items():
        scale = np.max(np.abs(ref))
        if scale == 0:
            assert float(grads[k].abs().max()) == 0, k
        else:
            assert float(np.max(np.abs(grads[k].numpy() - ref)) / scale) <= 1e-5, k
