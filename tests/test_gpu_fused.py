"""The fused inference path (gnnseg_fused.cu, hidden_dim 32 / 64), one kernel at a time against the oracle,
then the whole forward against the step-by-step kernels and the reference goldens.

State rows S (n + 1, 5h) = [SPs | SPd | Qi | Qo | Qs], SPs = exp(Ps) (= 2^(log2e Ps)), SPd = exp(Pd); row n is what an
absent neighbour reads (include/gnnseg.h).  Reference: gnn/model.py:69-81 (edge), :113-125 (node), :140-156.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_case, rel_err
from oracle import segclf_oracle as O
from test_gpu_parity import _steps_setup, make_model, sparse_graphs_of

pytestmark = pytest.mark.gpu
TOL = 1e-5
CASES = ["acts_ragged_h32_it4", "acts_h64_it6", "acts_masked_h32_it4"]


def state_rows(P, Q, h, extra=None):
    """Oracle projections -> n + 1 state rows, exponentials taken in fp64; the extra row is zeros, or `extra` (h floats:
    the start-node projection an absent start node reads in the final edge step) in its first h columns."""
    sp = torch.exp(P.double()).float()
    S = torch.cat([sp, Q], dim=1)
    last = torch.zeros(1, 5 * h)
    if extra is not None:
        last[0, :h] = extra
    return torch.cat([S, last], dim=0).contiguous()


def h1_oracle(p, H0, e, src, dst):
    """tanh(W3.[mi; mo; H] + b3): the first layer of the node network (gnn/model.py:114-122)."""
    real_i, real_o = dst >= 0, src >= 0
    mi = torch.zeros_like(H0).index_add_(0, dst[real_i], (e[:, None] * O._gather(H0, src))[real_i])
    mo = torch.zeros_like(H0).index_add_(0, src[real_o], (e[:, None] * O._gather(H0, dst))[real_o])
    return torch.tanh(O._lin(torch.cat([mi, mo, H0], dim=1), p, 3))


@pytest.mark.parametrize("name", CASES + ["half_edges_h8_it2"])
def test_fused_steps_against_oracle(name, cuda_device):
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    rec = load_case(name)
    model, batch, L = _steps_setup(rec, cuda_device)
    p = O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"])
    h, F, n = rec["h"], rec["F"], batch.n_nodes
    blob = model.pack_weights()
    st = _stream_ptr(cuda_device)
    src, dst, Xh = batch.src.cpu().long(), batch.dst.cpu().long(), batch.X.cpu()
    status = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    X4 = torch.full((n, 4), 7.0, device=cuda_device)
    S = torch.zeros(n + 1, 5 * h, device=cuda_device)
    rc = L.gnnseg_state_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(S), 5 * h, _ptr(status), st)
    if h not in (32, 64):
        assert rc == -2 and L.gnnseg_fused_gather_step(_ptr(blob), C.byref(batch.struct), _ptr(S), h, _ptr(S), 5 * h, st) == -2
        return
    assert rc == 0
    # ---- input step: state rows of [H0 | X]
    H0 = O.sparse_input(p, Xh)
    Pref, Qref = O.projections(p, H0)
    Sref = state_rows(Pref, Qref, h)
    assert torch.equal(X4.cpu()[:, :F], Xh) and torch.all(X4.cpu()[:, F:] == 0)
    assert torch.allclose(S.cpu(), Sref, rtol=1e-5, atol=3e-6) and torch.all(S[n] == 0)
    S2 = torch.full((n + 1, 5 * h), 9.0, device=cuda_device)
    assert L.gnnseg_state_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(S2), 2 * h, _ptr(status), st) == 0
    assert torch.allclose(S2.cpu()[:n, :2 * h], Sref[:n, :2 * h], rtol=1e-5, atol=3e-6)
    assert torch.all(S2[:, 2 * h:] == 9.0) and torch.all(S2[n] == 9.0)       # only the edge projections of the real rows
    # ---- fused edge + gather on the oracle's state rows: h1, in its own buffer and inside wider rows
    e_ref = O.sparse_edge(p, H0, src, dst)
    h1_ref = h1_oracle(p, H0, e_ref, src, dst)
    Sd = Sref.to(cuda_device)
    for ld in (h, 5 * h):
        h1 = torch.full((n, ld), 9.0, device=cuda_device)
        assert L.gnnseg_fused_gather_step(_ptr(blob), C.byref(batch.struct), _ptr(Sd), h, _ptr(h1), ld, st) == 0
        assert torch.allclose(h1.cpu()[:, :h], h1_ref, rtol=1e-5, atol=3e-6)
        assert torch.all(h1[:, h:] == 9.0)
    # ---- tensor-core MLP writing state rows; h1 inside the rows it becomes
    H1 = torch.cat([torch.tanh(O._lin(h1_ref, p, 4)), Xh], dim=1)
    P1ref, Q1ref = O.projections(p, H1)
    S1ref = state_rows(P1ref, Q1ref, h)
    h1d = h1_ref.to(cuda_device).contiguous()
    S1 = torch.zeros(n + 1, 5 * h, device=cuda_device)
    assert L.gnnseg_state_mlp_step(_ptr(blob), _ptr(X4), _ptr(h1d), h, n, h, _ptr(S1), 5 * h, _ptr(status), st) == 0
    assert torch.allclose(S1.cpu(), S1ref, rtol=1e-5, atol=4e-6)
    S1b = torch.zeros(n + 1, 5 * h, device=cuda_device)
    S1b[:n, :h] = h1d
    assert L.gnnseg_state_mlp_step(_ptr(blob), _ptr(X4), _ptr(S1b), 5 * h, n, h, _ptr(S1b), 5 * h, _ptr(status), st) == 0
    assert torch.equal(S1b, S1)
    S1c = torch.zeros(n + 1, 5 * h, device=cuda_device)                      # last step: only [SPs | SPd]
    S1c[:n, :h] = h1d
    assert L.gnnseg_state_mlp_step(_ptr(blob), _ptr(X4), _ptr(S1c), 5 * h, n, h, _ptr(S1c), 2 * h, _ptr(status), st) == 0
    assert torch.equal(S1c[:, :2 * h], S1[:, :2 * h]) and torch.all(S1c[:, 2 * h:] == 0)
    # ---- final edge step; an absent start node reads exp(b1) from the extra row
    e1_ref = O.sparse_edge(p, H1, src, dst)
    b1 = p[O.PARAM_KEYS[3]]
    S1d = state_rows(P1ref, Q1ref, h, extra=torch.exp(b1.double()).float()).to(cuda_device)
    sc = torch.full((batch.n_slots,), -1.0, device=cuda_device)
    assert L.gnnseg_edge_final_step(_ptr(blob), C.byref(batch.struct), _ptr(S1d), 5 * h, 0, h, h, _ptr(sc), st) == 0
    assert rel_err(sc.cpu().numpy(), e1_ref.numpy()) <= TOL
    assert int(status.item()) == 0


def test_adjacency_is_in_edges_then_out_edges(cuda_device):
    """gnnseg_build_adjacency: adj_ptr = in_ptr + out_ptr; the list of node v starts at 4*ceil(adj_ptr[v]/4) + 4v and holds
    the start nodes of its in-edges, then the end nodes of its out-edges with bit 31 set, in ascending slot order, padded
    to a multiple of four with n_nodes (which also stands for an absent neighbour); node_order lists every window's
    nodes by decreasing number of four-entry batches."""
    rec = load_case("half_edges_h8_it2")
    _, batch, _ = _steps_setup(rec, cuda_device)
    n = batch.n_nodes
    ip, op = batch.in_ptr.cpu().numpy().astype(np.int64), batch.out_ptr.cpu().numpy().astype(np.int64)
    inb, onb = batch.in_nbr.cpu().numpy().astype(np.int64), batch.out_nbr.cpu().numpy().astype(np.int64)
    ap, adj = batch.adj_ptr.cpu().numpy().astype(np.int64), batch.adj.cpu().numpy().view(np.uint32).astype(np.int64)
    assert np.array_equal(ap, ip + op)
    saw_absent = False
    trips = np.zeros(n, np.int64)
    for v in range(n):
        want = [x if x >= 0 else n for x in inb[ip[v]:ip[v + 1]]] + \
               [(x | 0x80000000) if x >= 0 else n for x in onb[op[v]:op[v + 1]]]
        saw_absent |= n in want
        trips[v] = (len(want) + 3) // 4
        want += [n] * (-len(want) % 4)
        off = 4 * ((ap[v] + 3) // 4) + 4 * v
        assert list(adj[off:off + len(want)]) == want
    assert saw_absent
    order = batch.node_order.cpu().numpy()[:n]
    assert sorted(order.tolist()) == list(range(n))
    assert np.all(np.diff(trips[order]) <= 0)                     # one window here: longest lists first


@pytest.mark.parametrize("name", CASES + ["toy2d_f2_h32_it10"])
def test_fused_forward_matches_reference_and_exact_kernels(name, cuda_device):
    """The whole forward: fused path (default) vs the reference's stored output and vs the step-by-step
    kernels (model.exact = True); reproducible to the bit; nothing clamped at these weights."""
    rec = load_case(name)
    model = make_model(rec, cuda_device)
    graphs = sparse_graphs_of(rec)
    with torch.no_grad():
        a = model(graphs).clone()
        b = model(graphs).clone()
        model.exact = True
        c = model(graphs).clone()
        model.exact = False
    model.check_range()
    assert torch.equal(a, b)
    assert rel_err(a.cpu().numpy(), rec["out"]) <= TOL
    assert rel_err(a.cpu().numpy(), c.cpu().numpy()) <= TOL
    assert not torch.equal(a, c)                                  # they are different kernels


def test_zero_iterations_and_empty_batches(cuda_device):
    from gnn_fpga_b200 import SegmentClassifier, SparseGraph, data
    g = data.acts_like_graph(30, seed=0)
    p = O.init_params(3, 32, seed=2)
    model = SegmentClassifier(3, 32, 0)
    model.load_state_dict(p)
    model = model.to(cuda_device).eval()
    X, src, dst, e_max = O.flatten_sparse_batch([g])
    with torch.no_grad():
        out = model([g])[0].cpu().numpy()
    assert rel_err(out, O.sparse_forward(p, X, src, dst, 0).numpy()) <= TOL
    empty = SparseGraph(np.zeros((5, 3), np.float32), *(np.zeros(0, np.int64),) * 4, np.zeros(0, np.float32))
    model2 = SegmentClassifier(3, 32, 2).to(cuda_device).eval()
    with torch.no_grad():
        assert tuple(model2([empty]).shape) == (1, 0)
        both = model2([g, empty])
        assert tuple(both.shape) == (2, e_max)


def test_range_flag_raises_and_exact_path_is_the_cure(cuda_device):
    """Projections beyond the fused path's range (|W1.[H|X] + b1| > 43.6) raise the flag: predict_stream refuses
    the batch, model.check_range() reports it, model.exact = True computes it with the step-by-step kernels."""
    from gnn_fpga_b200 import GnnsegError, SegmentClassifier, data
    g = data.acts_like_graph(40, seed=3)
    p = {k: (v * 300.0 if k.startswith("edge_network.network.0") else v) for k, v in O.init_params(3, 32, seed=0).items()}
    model = SegmentClassifier(3, 32, 2)
    model.load_state_dict(p)
    model = model.to(cuda_device).eval()
    with torch.no_grad():
        model([g])
        with pytest.raises(GnnsegError, match="range"):
            model.check_range()
        with pytest.raises(GnnsegError, match="range"):
            list(model.predict_stream([[g]]))
        model.exact = True
        out = model([g])[0].cpu().numpy()
        model.check_range()
    X, src, dst, e_max = O.flatten_sparse_batch([g])
    ref = O.sparse_forward(p, X, src, dst, 2, torch.float64).numpy()
    assert np.max(np.abs(out - ref)) <= 1e-4


_AB_SCRIPT = """
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
sys.path.insert(0, {tests!r})
from conftest import load_case
from test_gpu_parity import make_model, sparse_graphs_of
dev = torch.device("cuda:0")
out = {{}}
for name in {cases!r}:
    rec = load_case(name)
    model = make_model(rec, dev)
    with torch.no_grad():
        out[name] = model(sparse_graphs_of(rec)).cpu().numpy()
np.savez({dst!r}, **out)
"""


def test_pipeline_mlp_kernels_equal_the_round1_kernels(tmp_path, cuda_device):
    """The warp-specialised MLP / input kernels (gnnseg_mlp_pipe.cu, the default) against the two-CTA / serial tcgen05
    kernels of gnnseg_node_tc.cu (GNNSEG_MLP_PIPE=0): same GEMMs, same 3xTF32 split, same k order, so the forward's
    scores must agree to the last bit.  The switch is read once per process: two child processes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for pipe in ("1", "0"):
        dst = str(tmp_path / ("scores_pipe%s.npz" % pipe))
        code = _AB_SCRIPT.format(root=root, tests=os.path.join(root, "tests"), cases=CASES, dst=dst)
        env = dict(os.environ, GNNSEG_MLP_PIPE=pipe)
        subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
        res[pipe] = np.load(dst)
    for name in CASES:
        assert np.array_equal(res["1"][name], res["0"][name]), name
