"""Parity of the CUDA path (through the C ABI) with the reference goldens and the oracle.

Tolerances: indices / CSR bit-exact; edge scores <= 1e-5 relative in fp32
(BASELINE.json north_star).  Everything here needs a GPU.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import MODEL_CASES, load_case, rel_err
from oracle import segclf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def make_model(rec, device):
    from gnn_fpga_b200 import SegmentClassifier
    m = SegmentClassifier(rec["F"], rec["h"], rec["n_iters"], masks_e=rec["masks_e"], masks_n=rec["masks_n"])
    m.load_state_dict(rec["params"])
    return m.to(device).eval()


def sparse_graphs_of(rec):
    """Per-event SparseGraph tuples (np.nonzero of each event's unpadded block)."""
    from gnn_fpga_b200 import make_sparse_graph
    out = []
    for b in range(rec["X"].shape[0]):
        Ri, Ro = rec["Ri"][b], rec["Ro"][b]
        n_e = int(max(np.nonzero(Ri.sum(0))[0].max(initial=-1), np.nonzero(Ro.sum(0))[0].max(initial=-1)) + 1)
        n_n = int(max(np.nonzero(Ri.sum(1))[0].max(initial=-1), np.nonzero(Ro.sum(1))[0].max(initial=-1)) + 1)
        out.append(make_sparse_graph(rec["X"][b, :n_n], Ri[:n_n, :n_e], Ro[:n_n, :n_e], np.zeros(n_e, np.float32)))
    return out


# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MODEL_CASES)
def test_dense_call_matches_reference(name, cuda_device):
    """model([X, Ri, Ro]) -- the reference call -- against the reference's stored output."""
    rec = load_case(name)
    model = make_model(rec, cuda_device)
    X = torch.from_numpy(rec["X"]).to(cuda_device)
    Ri = torch.from_numpy(rec["Ri"].astype(np.float32)).to(cuda_device)
    Ro = torch.from_numpy(rec["Ro"].astype(np.float32)).to(cuda_device)
    with torch.no_grad():
        out = model([X, Ri, Ro])
    assert out.shape == rec["out"].shape and out.dtype == torch.float32
    assert rel_err(out.cpu().numpy(), rec["out"]) <= TOL


@pytest.mark.parametrize("name", ["acts_ragged_h32_it4", "acts_h64_it6", "c1_toy2d_h8_it1", "acts_masked_h8_it4"])
def test_sparse_batch_call_matches_reference(name, cuda_device):
    """List of SparseGraph tuples (no densifying) gives the same padded (B, E_max) scores."""
    rec = load_case(name)
    if name == "half_edges_h8_it2":
        pytest.skip("graph_from_sparse cannot express it")
    model = make_model(rec, cuda_device)
    graphs = sparse_graphs_of(rec)
    with torch.no_grad():
        out = model(graphs)
    assert tuple(out.shape) == rec["out"].shape
    assert rel_err(out.cpu().numpy(), rec["out"]) <= TOL


def test_masked_twin_bit_equal(cuda_device):
    """model.py and model_maskedlinear.py agree to the last bit in the reference; so do we."""
    a, b = load_case("acts_masked_h8_it4"), load_case("acts_masked_h8_it4_twin")
    outs = []
    for rec in (a, b):
        model = make_model(rec, cuda_device)
        outs.append(model(sparse_graphs_of(rec)).clone())
    assert torch.equal(outs[0], outs[1])
    # masks passed explicitly == weights zeroed beforehand and no masks
    rec = a
    from gnn_fpga_b200 import SegmentClassifier
    m2 = SegmentClassifier(rec["F"], rec["h"], rec["n_iters"])
    m2.load_state_dict(O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"]))
    m2 = m2.to(cuda_device).eval()
    assert torch.equal(m2(sparse_graphs_of(rec)), outs[0])


# ---------------------------------------------------------------------------------------
def test_dense_to_edges_and_csr_bit_exact(cuda_device):
    from gnn_fpga_b200 import DeviceGraphBatch
    for name in ("acts_ragged_h32_it4", "half_edges_h8_it2", "c1_toy2d_h8_it1"):
        rec = load_case(name)
        B, N, F = rec["X"].shape
        batch = DeviceGraphBatch.from_dense(torch.from_numpy(rec["X"]).to(cuda_device),
                                            torch.from_numpy(rec["Ri"].astype(np.float32)).to(cuda_device),
                                            torch.from_numpy(rec["Ro"].astype(np.float32)).to(cuda_device))
        src, dst = O.edges_from_dense(rec["Ri"], rec["Ro"])
        assert np.array_equal(batch.src.cpu().numpy(), src)
        assert np.array_equal(batch.dst.cpu().numpy(), dst)
        for key, other, ptr, eid, nbr in ((dst, src, batch.in_ptr, batch.in_eid, batch.in_nbr),
                                          (src, dst, batch.out_ptr, batch.out_eid, batch.out_nbr)):
            optr, oeid = O.csr_from_keys(key, B * N)
            assert np.array_equal(ptr.cpu().numpy(), optr)
            n_used = int(optr[-1])
            assert np.array_equal(eid.cpu().numpy()[:n_used], oeid)
            assert np.array_equal(nbr.cpu().numpy()[:n_used], other[oeid])
        for key, pos, eid in ((dst, batch.in_pos, batch.in_eid), (src, batch.out_pos, batch.out_eid)):
            pos, eid = pos.cpu().numpy(), eid.cpu().numpy()
            assert np.array_equal(pos < 0, key < 0)                       # absent endpoint <=> not in the CSR
            assert np.array_equal(eid[pos[key >= 0]], np.nonzero(key >= 0)[0])   # eid[pos[j]] == j


def test_csr_matches_reference_nonzero_golden(cuda_device):
    """Device CSR == np.nonzero tuples written by the reference's make_sparse_graph."""
    import os
    from conftest import GOLDEN
    from gnn_fpga_b200 import DeviceGraphBatch
    z = np.load(os.path.join(GOLDEN, "graph_roundtrip.npz"))
    batch = DeviceGraphBatch.from_dense(torch.from_numpy(z["X"][None]).to(cuda_device),
                                        torch.from_numpy(z["Ri"][None].astype(np.float32)).to(cuda_device),
                                        torch.from_numpy(z["Ro"][None].astype(np.float32)).to(cuda_device))
    n = z["Ri"].shape[0]
    in_ptr = batch.in_ptr.cpu().numpy()
    assert np.array_equal(batch.in_eid.cpu().numpy()[:in_ptr[-1]], z["Ri_cols"])
    assert np.array_equal(np.repeat(np.arange(n), np.diff(in_ptr)), z["Ri_rows"])
    out_ptr = batch.out_ptr.cpu().numpy()
    assert np.array_equal(batch.out_eid.cpu().numpy()[:out_ptr[-1]], z["Ro_cols"])
    assert np.array_equal(np.repeat(np.arange(n), np.diff(out_ptr)), z["Ro_rows"])


def test_csr_large_random_with_hub(cuda_device):
    """Unsorted endpoints, a degree-5000 hub (heapsort branch), degree-0 nodes, sentinels."""
    from gnn_fpga_b200 import DeviceGraphBatch
    rng = np.random.RandomState(0)
    n, m = 20000, 300000
    src = rng.randint(0, n, m)
    dst = rng.randint(0, n, m)
    dst[rng.choice(m, 5000, replace=False)] = 7
    src[rng.choice(m, 1000, replace=False)] = -1
    dst[rng.choice(m, 1000, replace=False)] = -1
    src[src == 11] = 12                                      # node 11 has no out-edges
    X = torch.zeros(n, 3, device=cuda_device)
    batch = DeviceGraphBatch(X, torch.from_numpy(src.astype(np.int32)).to(cuda_device),
                             torch.from_numpy(dst.astype(np.int32)).to(cuda_device), 1, m)
    for key, other, ptr, eid, nbr in ((dst, src, batch.in_ptr, batch.in_eid, batch.in_nbr),
                                      (src, dst, batch.out_ptr, batch.out_eid, batch.out_nbr)):
        optr, oeid = O.csr_from_keys(key, n)
        assert np.array_equal(ptr.cpu().numpy(), optr)
        assert np.array_equal(eid.cpu().numpy()[:optr[-1]], oeid)
        assert np.array_equal(nbr.cpu().numpy()[:optr[-1]], other[oeid])


def test_bad_incidence_raises(cuda_device):
    from gnn_fpga_b200 import DeviceGraphBatch
    X = torch.zeros(1, 4, 3, device=cuda_device)
    Ri = torch.zeros(1, 4, 5, device=cuda_device)
    Ro = torch.zeros(1, 4, 5, device=cuda_device)
    Ri[0, 1, 2] = 0.5
    with pytest.raises(ValueError, match="only 0 and 1"):
        DeviceGraphBatch.from_dense(X, Ri, Ro)
    Ri[0, 1, 2] = 1.0
    Ri[0, 3, 2] = 1.0
    with pytest.raises(ValueError, match="more than one"):
        DeviceGraphBatch.from_dense(X, Ri, Ro)


# ---------------------------------------------------------------------------------------
def _steps_setup(rec, device):
    from gnn_fpga_b200 import DeviceGraphBatch, _lib
    model = make_model(rec, device)
    B, N, F = rec["X"].shape
    batch = DeviceGraphBatch.from_dense(torch.from_numpy(rec["X"]).to(device),
                                        torch.from_numpy(rec["Ri"].astype(np.float32)).to(device),
                                        torch.from_numpy(rec["Ro"].astype(np.float32)).to(device))
    return model, batch, _lib.lib()


@pytest.mark.parametrize("name", ["acts_ragged_h32_it4", "acts_h64_it6", "toy2d_h4_it1", "toy2d_h16_it3", "acts_masked_h8_it4"])
def test_each_kernel_against_oracle(name, cuda_device):
    """gnnseg_input_step / gnnseg_edge_step / gnnseg_node_step one at a time vs the oracle.
    The state between kernels is X4, P = [W1a.HX+b1 | W1b.HX], Q = [W3a.HX | W3b.HX | W3c.HX+b3]
    (include/gnnseg.h); the oracle computes HX the reference's way and projects it."""
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    rec = load_case(name)
    model, batch, L = _steps_setup(rec, cuda_device)
    p = O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"])
    h, F, n = rec["h"], rec["F"], batch.n_nodes
    blob = model.pack_weights()
    X4 = torch.full((n, 4), 7.0, device=cuda_device)
    P = torch.zeros(n, 2 * h, device=cuda_device)
    Q = torch.zeros(n, 3 * h, device=cuda_device)
    P2 = torch.zeros_like(P)
    Q2 = torch.zeros_like(Q)
    e = torch.zeros(batch.n_slots, device=cuda_device)
    st = _stream_ptr(cuda_device)
    src = batch.src.cpu().long()
    dst = batch.dst.cpu().long()
    Xh = batch.X.cpu()

    assert L.gnnseg_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(P), _ptr(Q), st) == 0
    H0 = O.sparse_input(p, Xh)                                   # (n, h+F)
    Pref, Qref = O.projections(p, H0)
    assert torch.equal(X4.cpu()[:, :F], Xh) and torch.all(X4.cpu()[:, F:] == 0)
    assert torch.allclose(P.cpu(), Pref, rtol=1e-5, atol=2e-6)
    assert torch.allclose(Q.cpu(), Qref, rtol=1e-5, atol=2e-6)

    e_in = torch.full((batch.n_slots,), -1.0, device=cuda_device)
    e_out = torch.full((batch.n_slots,), -1.0, device=cuda_device)
    assert L.gnnseg_edge_step(_ptr(blob), C.byref(batch.struct), _ptr(P), h, _ptr(e), _ptr(e_in), _ptr(e_out), st) == 0
    e_ref = O.sparse_edge(p, H0, src, dst)
    assert rel_err(e.cpu().numpy(), e_ref.numpy()) <= TOL
    # the CSR-ordered copies hold exactly the same numbers: e_in[s] == e[in_eid[s]]
    n_in, n_out = int(batch.in_ptr[-1]), int(batch.out_ptr[-1])
    assert torch.equal(e_in[:n_in], e[batch.in_eid[:n_in].long()])
    assert torch.equal(e_out[:n_out], e[batch.out_eid[:n_out].long()])

    # node step on the oracle's own e and Q, so that only this kernel is under test
    e_dev = e_ref.to(cuda_device)
    e_in = e_dev[batch.in_eid[:n_in].long()].contiguous()
    e_out = e_dev[batch.out_eid[:n_out].long()].contiguous()
    Q_in = Qref.to(cuda_device).contiguous()
    assert L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(X4), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h,
                              _ptr(P2), _ptr(Q2), st) == 0
    H1 = torch.cat([O.sparse_node(p, H0, e_ref, src, dst), Xh], dim=1)
    P1ref, Q1ref = O.projections(p, H1)
    assert torch.allclose(P2.cpu(), P1ref, rtol=1e-5, atol=3e-6)
    assert torch.allclose(Q2.cpu(), Q1ref, rtol=1e-5, atol=3e-6)
    # Q_out = NULL (last iteration): P' identical, nothing else written
    P3 = torch.zeros_like(P)
    assert L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(X4), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h,
                              _ptr(P3), None, st) == 0
    assert torch.equal(P3, P2)


@pytest.mark.parametrize("name", ["acts_ragged_h32_it4", "acts_h64_it6", "toy2d_h16_it3", "half_edges_h8_it2"])
def test_node_step_halves_against_oracle(name, cuda_device):
    """gnnseg_node_gather_step (ordered CSR sum + tanh, every width) and gnnseg_node_mlp_step
    (tcgen05 layer 2 + projections, hidden_dim = 32 and 64) one at a time against the oracle, with h1 in
    its own buffer and aliased into the rows of P_out the way gnnseg_node_step passes it."""
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    rec = load_case(name)
    model, batch, L = _steps_setup(rec, cuda_device)
    p = O.apply_masks(rec["params"], rec["masks_e"], rec["masks_n"])
    h, F, n = rec["h"], rec["F"], batch.n_nodes
    blob = model.pack_weights()
    st = _stream_ptr(cuda_device)
    src, dst, Xh = batch.src.cpu().long(), batch.dst.cpu().long(), batch.X.cpu()
    H0 = O.sparse_input(p, Xh)
    _, Qref = O.projections(p, H0)
    e_ref = O.sparse_edge(p, H0, src, dst)
    n_in, n_out = int(batch.in_ptr[-1]), int(batch.out_ptr[-1])
    e_dev = e_ref.to(cuda_device)
    e_in = e_dev[batch.in_eid[:n_in].long()].contiguous()
    e_out = e_dev[batch.out_eid[:n_out].long()].contiguous()
    Q_in = Qref.to(cuda_device).contiguous()
    # oracle h1 = tanh(W3.[mi; mo; H] + b3), the first layer of the node network (gnn/model.py:114-122)
    real_i, real_o = dst >= 0, src >= 0
    mi = torch.zeros_like(H0).index_add_(0, dst[real_i], (e_ref[:, None] * O._gather(H0, src))[real_i])
    mo = torch.zeros_like(H0).index_add_(0, src[real_o], (e_ref[:, None] * O._gather(H0, dst))[real_o])
    h1_ref = torch.tanh(O._lin(torch.cat([mi, mo, H0], dim=1), p, 3))
    for ld in (h, 2 * h):
        h1 = torch.full((n, ld), 9.0, device=cuda_device)
        assert L.gnnseg_node_gather_step(C.byref(batch.struct), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h, _ptr(h1), ld, st) == 0
        assert torch.allclose(h1.cpu()[:, :h], h1_ref, rtol=1e-5, atol=2e-6)
        assert torch.all(h1[:, h:] == 9.0)                      # nothing written beyond the h columns
    assert L.gnnseg_node_gather_step(C.byref(batch.struct), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h, _ptr(h1), h - 1, st) == -1
    X4 = torch.zeros(n, 4, device=cuda_device)
    X4[:, :F] = batch.X
    P2, Q2 = torch.zeros(n, 2 * h, device=cuda_device), torch.zeros(n, 3 * h, device=cuda_device)
    h1d = h1_ref.to(cuda_device).contiguous()
    rc = L.gnnseg_node_mlp_step(_ptr(blob), _ptr(X4), _ptr(h1d), h, n, h, _ptr(P2), _ptr(Q2), st)
    if h not in (32, 64):                                        # the tcgen05 MLP kernels; other widths run the fused kernel
        assert rc == -2
        return
    assert rc == 0
    H1 = torch.cat([torch.tanh(O._lin(h1_ref, p, 4)), Xh], dim=1)
    P1ref, Q1ref = O.projections(p, H1)
    assert torch.allclose(P2.cpu(), P1ref, rtol=1e-5, atol=3e-6)
    assert torch.allclose(Q2.cpu(), Q1ref, rtol=1e-5, atol=3e-6)
    # h1 inside the rows of P_out (stride 2h): same result to the bit
    P3, Q3 = torch.zeros_like(P2), torch.zeros_like(Q2)
    P3[:, :h] = h1d
    assert L.gnnseg_node_mlp_step(_ptr(blob), _ptr(X4), _ptr(P3), 2 * h, n, h, _ptr(P3), _ptr(Q3), st) == 0
    assert torch.equal(P3, P2) and torch.equal(Q3, Q2)


def test_fused_and_split_node_step_agree(cuda_device):
    """hidden_dim = 32: gnnseg_node_step (gather kernel + tcgen05 MLP kernel) and the two halves
    called one after the other give the same bits."""
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    rec = load_case("acts_ragged_h32_it4")
    model, batch, L = _steps_setup(rec, cuda_device)
    h, F, n = rec["h"], rec["F"], batch.n_nodes
    blob = model.pack_weights()
    st = _stream_ptr(cuda_device)
    g = torch.Generator(cuda_device).manual_seed(1)
    X4 = torch.zeros(n, 4, device=cuda_device)
    X4[:, :F] = batch.X
    Q_in = torch.randn(n, 3 * h, device=cuda_device, generator=g) * 0.3
    e_in = torch.rand(batch.n_slots, device=cuda_device, generator=g)
    e_out = torch.rand(batch.n_slots, device=cuda_device, generator=g)
    Pa, Qa = torch.zeros(n, 2 * h, device=cuda_device), torch.zeros(n, 3 * h, device=cuda_device)
    assert L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(X4), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h,
                              _ptr(Pa), _ptr(Qa), st) == 0
    h1 = torch.zeros(n, h, device=cuda_device)
    Pb, Qb = torch.zeros_like(Pa), torch.zeros_like(Qa)
    assert L.gnnseg_node_gather_step(C.byref(batch.struct), _ptr(Q_in), _ptr(e_in), _ptr(e_out), h, _ptr(h1), h, st) == 0
    assert L.gnnseg_node_mlp_step(_ptr(blob), _ptr(X4), _ptr(h1), h, n, h, _ptr(Pb), _ptr(Qb), st) == 0
    assert torch.equal(Pa, Pb) and torch.equal(Qa, Qb)


@pytest.mark.parametrize("name,impl", [("acts_ragged_h32_it4", "fused"), ("acts_ragged_h32_it4", "mma"), ("acts_h64_it6", "mma")])
def test_alternative_node_step_implementations(name, impl, cuda_device, monkeypatch):
    """The kernels kept for A/B runs (GNNSEG_NODE_IMPL=fused: one tcgen05 kernel with the gather inside;
    =mma: the generic fused kernel on mma.sync 3xTF32) still meet the same gate as the default path."""
    rec = load_case(name)
    model = make_model(rec, cuda_device)
    model.use_cuda_graph = False
    graphs = sparse_graphs_of(rec)
    with torch.no_grad():
        ref = model(graphs).clone()
        monkeypatch.setenv("GNNSEG_NODE_IMPL", impl)
        alt = model(graphs).clone()
    assert rel_err(alt.cpu().numpy(), rec["out"]) <= TOL
    assert rel_err(alt.cpu().numpy(), ref.cpu().numpy()) <= 2e-6


def test_deterministic_and_graph_replay(cuda_device):
    """Same bits run to run, with and without the CUDA graph, and after rebuilding the batch."""
    from gnn_fpga_b200 import DeviceGraphBatch
    rec = load_case("acts_ragged_h32_it4")
    model = make_model(rec, cuda_device)
    graphs = sparse_graphs_of(rec)
    batch = DeviceGraphBatch.from_sparse_graphs(graphs, cuda_device)
    model.use_cuda_graph = False
    a = model(batch).clone()
    model.use_cuda_graph = True
    b = model(batch).clone()      # warm
    c = model(batch).clone()      # capture + replay
    d = model(batch).clone()      # replay
    e = model(DeviceGraphBatch.from_sparse_graphs(graphs, cuda_device)).clone()
    for t in (b, c, d, e):
        assert torch.equal(a, t)
    # replay picks up changed weights (the blob is re-packed in place before each replay)
    with torch.no_grad():
        model.edge_network.network[2].bias.add_(0.25)
    f = model(batch).clone()
    model.use_cuda_graph = False
    g = model(batch).clone()
    assert torch.equal(f, g) and not torch.equal(f, a)


def test_events_are_independent(cuda_device):
    """Scores of an event do not depend on what else is in the batch: batch == one by one
    (bit-equal), which is also what makes the multi-GPU event sharding exact."""
    from gnn_fpga_b200 import data, SegmentClassifier
    torch.manual_seed(0)
    model = SegmentClassifier(3, 32, 4).to(cuda_device).eval()
    graphs = [data.acts_like_graph(n, seed=i) for i, n in enumerate((40, 55, 33, 47))]
    full = model(graphs).clone()
    for b, g in enumerate(graphs):
        one = model([g])
        n_e = g.Ri_rows.shape[0]
        assert torch.equal(one[0, :n_e], full[b, :n_e])


@pytest.mark.parametrize("h,n_iters", [(32, 3), (64, 3), (16, 2)])
def test_many_tiles_per_cta_events_independent(h, n_iters, cuda_device):
    """A batch large enough that every persistent CTA walks several 128-node tiles (44 k nodes =
    344 tiles on 148 SMs: tile-to-tile barriers, operand buffer reuse, weight halves re-streamed at
    h = 64): every event scores bit-equal to the same event run alone, and one event matches the oracle."""
    from gnn_fpga_b200 import data, SegmentClassifier
    graphs = [data.acts_like_graph(400, seed=100 + i) for i in range(11)]
    p = O.init_params(3, h, seed=4)
    model = SegmentClassifier(3, h, n_iters)
    model.load_state_dict(p)
    model = model.to(cuda_device).eval()
    with torch.no_grad():
        full = model(graphs).clone()
        for b in (0, 5, 10):
            one = model([graphs[b]])
            n_e = graphs[b].Ri_rows.shape[0]
            assert torch.equal(one[0, :n_e], full[b, :n_e])
    X, src, dst, e_max = O.flatten_sparse_batch([graphs[10]])
    ref = O.sparse_forward(p, X, src, dst, n_iters).numpy()
    assert rel_err(full[10, :e_max].cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("h,n_iters,n_tracks", [(32, 4, 400), (64, 8, 1000)])
def test_event_scale_against_sparse_oracle(h, n_iters, n_tracks, cuda_device):
    """One ACTS-like event (BASELINE configs[1] event size; a 1/10 mu200 event) vs the fp32 and
    fp64 sparse restatements: <= 1e-5 relative."""
    from gnn_fpga_b200 import data, SegmentClassifier
    g = data.acts_like_graph(n_tracks, seed=3)
    p = O.init_params(3, h, seed=0)
    model = SegmentClassifier(3, h, n_iters)
    model.load_state_dict(p)
    model = model.to(cuda_device).eval()
    out = model([g])[0].cpu().numpy()
    X, src, dst, e_max = O.flatten_sparse_batch([g])
    ref32 = O.sparse_forward(p, X, src, dst, n_iters, torch.float32).numpy()
    ref64 = O.sparse_forward(p, X, src, dst, n_iters, torch.float64).numpy()
    assert rel_err(out, ref32) <= TOL
    assert rel_err(out, ref64) <= TOL


def test_full_size_properties(cuda_device):
    """BASELINE configs[1] at full size (64 events, ~1.28M edges): scores in (0,1), padding
    slots equal the padding constant, and relabelling the nodes of every event (a permutation
    of node ids, edges kept in place) leaves every score within fp32 reordering noise."""
    from gnn_fpga_b200 import data, SegmentClassifier, SparseGraph
    from gnn_fpga_b200.data import sparse_from_edges
    graphs = data.acts_like_graphs(64, 400, seed=0)
    torch.manual_seed(0)
    model = SegmentClassifier(3, 32, 4).to(cuda_device).eval()
    out = model(graphs).clone()
    assert torch.all((out > 0) & (out < 1))
    p = {k: v.cpu() for k, v in model.state_dict().items()}
    const = torch.sigmoid(p[O.PARAM_KEYS[4]] @ torch.tanh(p[O.PARAM_KEYS[3]]) + p[O.PARAM_KEYS[5]]).item()
    for b, g in enumerate(graphs):
        pad = out[b, g.Ri_rows.shape[0]:].cpu().numpy()
        assert np.all(np.abs(pad - const) <= 1e-6 * const)
    rng = np.random.RandomState(5)
    permuted = []
    for g in graphs[:8]:
        n = g.X.shape[0]
        perm = rng.permutation(n)                 # new id of old node i
        src = np.empty(g.Ro_cols.shape[0], np.int64); src[g.Ro_cols] = g.Ro_rows
        dst = np.empty(g.Ri_cols.shape[0], np.int64); dst[g.Ri_cols] = g.Ri_rows
        Xp = np.empty_like(g.X); Xp[perm] = g.X
        permuted.append(sparse_from_edges(Xp, perm[src], perm[dst], g.y))
    out_p = model(permuted)
    for b, g in enumerate(graphs[:8]):
        n_e = g.Ri_rows.shape[0]
        assert rel_err(out_p[b, :n_e].cpu().numpy(), out[b, :n_e].cpu().numpy()) <= TOL


def test_unsupported_shapes_fail_loudly(cuda_device):
    from gnn_fpga_b200 import SegmentClassifier, GnnsegError, data
    for mode in ("eval", "train"):        # the CUDA path and the differentiable training path agree on this
        model = getattr(SegmentClassifier(3, 12, 1).to(cuda_device), mode)()
        with pytest.raises(GnnsegError, match="unsupported"):
            model(data.toy2d_graphs(1))
        model = getattr(SegmentClassifier(2, 8, 1).to(cuda_device), mode)()
        with pytest.raises(ValueError, match="input_dim"):
            model(data.toy2d_graphs(1, input_dim=3))


def test_empty_and_tiny_inputs(cuda_device):
    from gnn_fpga_b200 import SegmentClassifier, SparseGraph
    torch.manual_seed(1)
    model = SegmentClassifier(3, 8, 2).to(cuda_device).eval()
    # a graph with nodes but no edges, next to a normal one
    i64 = np.zeros(0, np.int64)
    g0 = SparseGraph(np.random.RandomState(0).rand(5, 3).astype(np.float32), i64, i64, i64, i64, np.zeros(0, np.float32))
    g1 = SparseGraph(np.random.RandomState(1).rand(3, 3).astype(np.float32), np.array([1, 2]), np.array([0, 1]),
                     np.array([0, 1]), np.array([0, 1]), np.zeros(2, np.float32))
    out = model([g0, g1])
    assert tuple(out.shape) == (2, 2)
    p = {k: v.cpu() for k, v in model.state_dict().items()}
    X, src, dst, e_max = O.flatten_sparse_batch([g0, g1])
    ref = O.sparse_forward(p, X, src, dst, 2).numpy().reshape(2, 2)
    assert rel_err(out.cpu().numpy(), ref) <= TOL
    # zero edges in the whole batch
    out = model([g0])
    assert tuple(out.shape) == (1, 0)


def test_predict_stream_matches_blocking_calls(cuda_device):
    """Pipelined host-to-host inference returns, in order, the same bits as model(graphs)."""
    from gnn_fpga_b200 import data, SegmentClassifier
    torch.manual_seed(3)
    model = SegmentClassifier(3, 32, 2).to(cuda_device).eval()
    batches = [[data.acts_like_graph(n, seed=10 * b + i) for i, n in enumerate((30 + b, 45, 25 + 2 * b))] for b in range(5)]
    expect = [model(g).cpu() for g in batches]
    got = [t.clone() for t in model.predict_stream(batches, depth=2)]
    assert len(got) == len(expect)
    for a, b in zip(got, expect):
        assert a.shape == b.shape and torch.equal(a, b)
    assert [t.shape for t in model.predict_stream(batches[:1], depth=3)] == [expect[0].shape]
    assert list(model.predict_stream([])) == []


def test_graph_files_match_tuples(cuda_device, tmp_path):
    """model(list of .npz file names) -- mapped, parsed and packed by the library -- gives the same
    bits as model(list of SparseGraph loaded with np.load), in the blocking call and in predict_stream."""
    from gnn_fpga_b200 import SegmentClassifier, SparseGraph, data, load_graphs, save_graphs
    graphs = [data.acts_like_graph(n, seed=20 + i) for i, n in enumerate((40, 25, 33, 12, 50, 8))]
    names = [str(tmp_path / ("event%06i.npz" % i)) for i in range(len(graphs))]
    save_graphs(graphs, names)
    torch.manual_seed(2)
    model = SegmentClassifier(3, 32, 3).to(cuda_device).eval()
    with torch.no_grad():
        a = model(load_graphs(names, SparseGraph)).clone()
        b = model(names).clone()
    assert a.shape == b.shape and torch.equal(a, b)
    batches = [names[:3], names[3:]]
    with torch.no_grad():
        expect = [model(load_graphs(bt, SparseGraph)).cpu() for bt in batches]
    got = [t.clone() for t in model.predict_stream(batches)]
    for x, y in zip(got, expect):
        assert torch.equal(x, y)


def test_single_direction_build_csr_entry_point(cuda_device):
    """gnnseg_build_csr (one key array) gives the same CSR as the paired gnnseg_build_graph."""
    from gnn_fpga_b200 import _lib, DeviceGraphBatch, data
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    L = _lib.lib()
    batch = DeviceGraphBatch.from_sparse_graphs([data.acts_like_graph(30, seed=1), data.acts_like_graph(22, seed=2)], cuda_device)
    n, m = batch.n_nodes, batch.n_slots
    i32 = dict(dtype=torch.int32, device=cuda_device)
    ptr, eid, nbr, pos = torch.empty(n + 1, **i32), torch.empty(m, **i32), torch.empty(m, **i32), torch.empty(m, **i32)
    wsb = L.gnnseg_csr_workspace_bytes(n, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=cuda_device)
    assert L.gnnseg_build_csr(_ptr(batch.dst), _ptr(batch.src), m, n, _ptr(ptr), _ptr(eid), _ptr(nbr), _ptr(pos),
                              _ptr(ws), wsb, _stream_ptr(cuda_device)) == 0
    k = int(ptr[-1])
    assert torch.equal(ptr, batch.in_ptr) and torch.equal(eid[:k], batch.in_eid[:k])
    assert torch.equal(nbr[:k], batch.in_nbr[:k]) and torch.equal(pos, batch.in_pos)
    assert L.gnnseg_build_csr(_ptr(batch.dst), _ptr(batch.src), m, n, _ptr(ptr), _ptr(eid), _ptr(nbr), None,
                              _ptr(ws), wsb // 2, _stream_ptr(cuda_device)) == -3      # workspace too small
