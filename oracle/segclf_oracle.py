"""CPU oracle for the SegmentClassifier hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; it is the checker, never the product (gnn_fpga_b200 has no CPU path).

It restates the algorithm of the reference in plain torch-CPU / numpy, as pure functions over
a dict of the ten parameter tensors (keys = the reference's state_dict keys):

  dense_forward   follows gnn/model.py:140-156 operation by operation (incidence bmm,
                  broadcast multiply, cat, linear), i.e. what the reference executes.
  sparse_forward  the same mathematics with the bmm against 0/1 matrices written as row
                  gathers and an ordered index_add (a bmm with a 0/1 matrix
                  is a gather; the segment sum differs only in summation order).  Used where
                  dense matrices cannot exist (41 GB / 800 GB at BASELINE configs 2 and 4).
  edges_from_dense / csr_from_keys   the integer side: np.nonzero order of gnn/graph.py:23-26.

PINNING: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports
/root/reference/gnn/{model,model_maskedlinear,graph}.py in the build container and stores
inputs, state_dicts and outputs under tests/golden/; tests/test_oracle_golden.py checks both
restatements against them (dense: bit-exact; sparse: <= 2e-6 relative).
"""
import numpy as np
import torch

PARAM_KEYS = (
    "input_network.0.weight", "input_network.0.bias",
    "edge_network.network.0.weight", "edge_network.network.0.bias",
    "edge_network.network.2.weight", "edge_network.network.2.bias",
    "node_network.network.0.weight", "node_network.network.0.bias",
    "node_network.network.2.weight", "node_network.network.2.bias",
)


def init_params(input_dim, hidden_dim, seed=0):
    """Default nn.Linear initialisation in the reference's construction order
    (gnn/model.py:132-138): input Linear, edge layers 0/2, node layers 0/2."""
    torch.manual_seed(seed)
    D = input_dim + hidden_dim
    layers = [torch.nn.Linear(input_dim, hidden_dim), torch.nn.Linear(2 * D, hidden_dim),
              torch.nn.Linear(hidden_dim, 1), torch.nn.Linear(3 * D, hidden_dim),
              torch.nn.Linear(hidden_dim, hidden_dim)]
    out = {}
    for i, lin in enumerate(layers):
        out[PARAM_KEYS[2 * i]] = lin.weight.detach().clone()
        out[PARAM_KEYS[2 * i + 1]] = lin.bias.detach().clone()
    return out


def apply_masks(params, masks_e=None, masks_n=None):
    """MaskedLinear: effective weight = weight * mask (gnn/model.py:28-31)."""
    p = dict(params)
    if masks_e is not None:
        p[PARAM_KEYS[2]] = p[PARAM_KEYS[2]] * masks_e[0]
        p[PARAM_KEYS[4]] = p[PARAM_KEYS[4]] * masks_e[1]
    if masks_n is not None:
        p[PARAM_KEYS[6]] = p[PARAM_KEYS[6]] * masks_n[0]
        p[PARAM_KEYS[8]] = p[PARAM_KEYS[8]] * masks_n[1]
    return p


def _lin(x, p, i):
    return torch.nn.functional.linear(x, p[PARAM_KEYS[2 * i]], p[PARAM_KEYS[2 * i + 1]])


# ---- dense restatement (what the reference runs) ---------------------------------------
def dense_edge(p, H, Ri, Ro):
    # gnn/model.py:71-81
    bo = torch.bmm(Ro.transpose(1, 2), H)
    bi = torch.bmm(Ri.transpose(1, 2), H)
    z = torch.tanh(_lin(torch.cat([bo, bi], dim=2), p, 1))
    return torch.sigmoid(_lin(z, p, 2)).squeeze(-1)


def dense_node(p, H, e, Ri, Ro):
    # gnn/model.py:114-125
    bo = torch.bmm(Ro.transpose(1, 2), H)
    bi = torch.bmm(Ri.transpose(1, 2), H)
    mi = torch.bmm(Ri * e[:, None], bo)
    mo = torch.bmm(Ro * e[:, None], bi)
    z = torch.tanh(_lin(torch.cat([mi, mo, H], dim=2), p, 3))
    return torch.tanh(_lin(z, p, 4))


def dense_forward(p, X, Ri, Ro, n_iters):
    # gnn/model.py:140-156
    with torch.no_grad():
        H = torch.cat([torch.tanh(_lin(X, p, 0)), X], dim=-1)
        for _ in range(n_iters):
            e = dense_edge(p, H, Ri, Ro)
            H = torch.cat([dense_node(p, H, e, Ri, Ro), X], dim=-1)
        return dense_edge(p, H, Ri, Ro)


# ---- sparse restatement ------------------------------------------------------------------
def _gather(H, idx):
    """Rows H[idx]; idx == -1 (absent endpoint, a zero column of Ri/Ro) gives the zero row."""
    out = H[idx.clamp(min=0)]
    return out * (idx >= 0).to(H.dtype)[:, None]


def sparse_input(p, X):
    return torch.cat([torch.tanh(_lin(X, p, 0)), X], dim=-1)


def sparse_edge(p, H, src, dst):
    z = torch.tanh(_lin(torch.cat([_gather(H, src), _gather(H, dst)], dim=1), p, 1))
    return torch.sigmoid(_lin(z, p, 2)).squeeze(-1)


def sparse_node(p, H, e, src, dst):
    n = H.shape[0]
    real_i = dst >= 0
    real_o = src >= 0
    mi = torch.zeros_like(H).index_add_(0, dst[real_i], (e[:, None] * _gather(H, src))[real_i])
    mo = torch.zeros_like(H).index_add_(0, src[real_o], (e[:, None] * _gather(H, dst))[real_o])
    z = torch.tanh(_lin(torch.cat([mi, mo, H], dim=1), p, 3))
    assert z.shape[0] == n
    return torch.tanh(_lin(z, p, 4))


def sparse_forward(p, X, src, dst, n_iters, dtype=torch.float32, return_state=False):
    """X (n_nodes,F); src/dst (n_slots,) int64 node ids, -1 = absent.  Returns (n_slots,)."""
    with torch.no_grad():
        p = {k: v.to(dtype) for k, v in p.items()}
        X = torch.as_tensor(X).to(dtype)
        src = torch.as_tensor(np.asarray(src)).long()
        dst = torch.as_tensor(np.asarray(dst)).long()
        H = sparse_input(p, X)
        for _ in range(n_iters):
            e = sparse_edge(p, H, src, dst)
            H = torch.cat([sparse_node(p, H, e, src, dst), X], dim=-1)
        out = sparse_edge(p, H, src, dst)
        return (out, H) if return_state else out


def sparse_vjp(p, X, src, dst, n_iters, dscores=None, y=None, l1=0.0, masks_e=None, masks_n=None,
               dtype=torch.float64):
    """Gradients of the sparse restatement by torch autograd on the CPU (the checker of
    gnnseg_backward).  Either a cotangent `dscores` (n_slots,) is given, or targets `y`: then the
    loss is Estimator.training_step's (gnn/estimator.py:53-57): BCELoss mean over ALL slots of the
    padded batch + l1 * sum|W| over the edge / node network weights.  Masks multiply the weights
    inside the graph (MaskedLinear, gnn/model.py:28-31), so dW = dW_eff * mask.
    Returns (scores, loss or None, {key: grad})."""
    leaf = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}
    q = dict(leaf)
    if masks_e is not None:
        q[PARAM_KEYS[2]] = leaf[PARAM_KEYS[2]] * masks_e[0].to(dtype)
        q[PARAM_KEYS[4]] = leaf[PARAM_KEYS[4]] * masks_e[1].to(dtype)
    if masks_n is not None:
        q[PARAM_KEYS[6]] = leaf[PARAM_KEYS[6]] * masks_n[0].to(dtype)
        q[PARAM_KEYS[8]] = leaf[PARAM_KEYS[8]] * masks_n[1].to(dtype)
    X = torch.as_tensor(X).to(dtype)
    src = torch.as_tensor(np.asarray(src)).long()
    dst = torch.as_tensor(np.asarray(dst)).long()
    H = sparse_input(q, X)
    for _ in range(n_iters):
        e = sparse_edge(q, H, src, dst)
        H = torch.cat([sparse_node(q, H, e, src, dst), X], dim=-1)
    out = sparse_edge(q, H, src, dst)
    loss = None
    if dscores is not None:
        out.backward(torch.as_tensor(dscores).to(dtype))
    else:
        loss = torch.nn.functional.binary_cross_entropy(out, torch.as_tensor(y).to(dtype).reshape(-1))
        if l1:
            loss = loss + l1 * sum(leaf[k].abs().sum() for k in (PARAM_KEYS[6], PARAM_KEYS[8], PARAM_KEYS[2], PARAM_KEYS[4]))
        loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return out.detach(), (loss.detach() if loss is not None else None), grads


# ---- NodeClassifier (gnn/MPNN_HitClassifier.ipynb cell 21) -------------------------------
# The body is the segment classifier's (cells 20-21 restate gnn/model.py's EdgeNetwork / NodeNetwork
# without masks); the head is output_network = Linear(D, 1) + Sigmoid on [H | X] per node, and there
# is no final edge step.  Pinned by tests/golden/nodeclf_*.npz (oracle/make_golden_nodeclf.py
# executes the notebook's own class definitions).
HEAD_KEYS = ("output_network.0.weight", "output_network.0.bias")


def _head(p, H):
    return torch.sigmoid(torch.nn.functional.linear(H, p[HEAD_KEYS[0]], p[HEAD_KEYS[1]])).squeeze(-1)


def nodeclf_dense_forward(p, X, Ri, Ro, n_iters):
    with torch.no_grad():
        H = torch.cat([torch.tanh(_lin(X, p, 0)), X], dim=-1)
        for _ in range(n_iters):
            e = dense_edge(p, H, Ri, Ro)
            H = torch.cat([dense_node(p, H, e, Ri, Ro), X], dim=-1)
        return _head(p, H)


def nodeclf_sparse_forward(p, X, src, dst, n_iters, dtype=torch.float32):
    """X (n_nodes, F); src/dst (n_slots,) node ids or -1.  Returns (n_nodes,) node scores."""
    with torch.no_grad():
        p = {k: v.to(dtype) for k, v in p.items()}
        X = torch.as_tensor(X).to(dtype)
        src = torch.as_tensor(np.asarray(src)).long()
        dst = torch.as_tensor(np.asarray(dst)).long()
        H = sparse_input(p, X)
        for _ in range(n_iters):
            e = sparse_edge(p, H, src, dst)
            H = torch.cat([sparse_node(p, H, e, src, dst), X], dim=-1)
        return _head(p, H)


def nodeclf_sparse_vjp(p, X, src, dst, n_iters, dnode=None, y=None, dtype=torch.float64):
    """Gradients of the node classifier's sparse restatement (checker of gnnseg_backward_nodes):
    cotangent `dnode` (n_nodes,), or targets `y` with nn.BCELoss() (mean over all nodes).
    Returns (scores, loss or None, {key: grad}); parameters the loss does not reach get zeros."""
    leaf = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}
    X = torch.as_tensor(X).to(dtype)
    src = torch.as_tensor(np.asarray(src)).long()
    dst = torch.as_tensor(np.asarray(dst)).long()
    H = sparse_input(leaf, X)
    for _ in range(n_iters):
        e = sparse_edge(leaf, H, src, dst)
        H = torch.cat([sparse_node(leaf, H, e, src, dst), X], dim=-1)
    out = _head(leaf, H)
    loss = None
    if dnode is not None:
        out.backward(torch.as_tensor(dnode).to(dtype))
    else:
        loss = torch.nn.functional.binary_cross_entropy(out, torch.as_tensor(y).to(dtype).reshape(-1))
        loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return out.detach(), (loss.detach() if loss is not None else None), grads


def projections(p, HX):
    """Per-node first-layer projections the CUDA path carries between kernels (same algebra as
    gnn/model.py:73-81,120-125: W.[a;b;c] = Wa.a + Wb.b + Wc.c).  HX (n, D) = [H | X].
    Returns P (n, 2h) = [W1a.HX + b1 | W1b.HX] and Q (n, 3h) = [W3a.HX | W3b.HX | W3c.HX + b3]."""
    D = HX.shape[1]
    W1, b1 = p[PARAM_KEYS[2]], p[PARAM_KEYS[3]]
    W3, b3 = p[PARAM_KEYS[6]], p[PARAM_KEYS[7]]
    P = torch.cat([HX @ W1[:, :D].T + b1, HX @ W1[:, D:].T], dim=1)
    Q = torch.cat([HX @ W3[:, :D].T, HX @ W3[:, D:2 * D].T, HX @ W3[:, 2 * D:].T + b3], dim=1)
    return P, Q


# ---- integer side --------------------------------------------------------------------------
def edges_from_dense(Ri, Ro):
    """(B,N,E) 0/1 arrays -> per-slot endpoints src/dst (B*E,), flattened node ids b*N+n,
    -1 where the column is empty.  Via np.nonzero as make_sparse_graph (gnn/graph.py:23-26)."""
    Ri = np.asarray(Ri)
    Ro = np.asarray(Ro)
    B, N, E = Ri.shape
    src = np.full(B * E, -1, dtype=np.int64)
    dst = np.full(B * E, -1, dtype=np.int64)
    b, n, e = np.nonzero(Ri)
    dst[b * E + e] = b * N + n
    b, n, e = np.nonzero(Ro)
    src[b * E + e] = b * N + n
    return src, dst


def csr_from_keys(key, n_nodes):
    """CSR of slots grouped by key in np.nonzero order: rows ascending, slot ids ascending
    within a row; key < 0 dropped.  rowptr = [0] + cumsum(bincount)."""
    key = np.asarray(key, dtype=np.int64)
    slot = np.nonzero(key >= 0)[0]
    order = np.lexsort((slot, key[slot]))
    eid = slot[order]
    ptr = np.concatenate([[0], np.cumsum(np.bincount(key[slot], minlength=n_nodes))])
    return ptr.astype(np.int64), eid.astype(np.int64)


def flatten_sparse_batch(graphs):
    """List of SparseGraph -> (X (Nt,F), src, dst (B*e_max,), e_max), padded the way
    graph_from_sparse + merge_graphs pad edges (gnn/graph.py:28-35,
    gnn/trainSegmentClassifier.py:66-95); node rows are not padded."""
    e_max = max(int(g.Ri_rows.shape[0]) for g in graphs)
    B = len(graphs)
    src = np.full(B * e_max, -1, dtype=np.int64)
    dst = np.full(B * e_max, -1, dtype=np.int64)
    off = 0
    Xs = []
    for b, g in enumerate(graphs):
        dst[b * e_max + np.asarray(g.Ri_cols)] = np.asarray(g.Ri_rows) + off
        src[b * e_max + np.asarray(g.Ro_cols)] = np.asarray(g.Ro_rows) + off
        Xs.append(np.asarray(g.X, dtype=np.float32))
        off += g.X.shape[0]
    return np.concatenate(Xs, axis=0), src, dst, e_max


def merge_dense(graphs):
    """merge_graphs (gnn/trainSegmentClassifier.py:66-95) on dense Graph tuples, fp32 out."""
    B = len(graphs)
    F = graphs[0].X.shape[1]
    n_max = max(g.X.shape[0] for g in graphs)
    e_max = max(g.Ri.shape[1] for g in graphs)
    X = np.zeros((B, n_max, F), np.float32)
    Ri = np.zeros((B, n_max, e_max), np.float32)
    Ro = np.zeros((B, n_max, e_max), np.float32)
    for b, g in enumerate(graphs):
        n, e = g.Ri.shape
        X[b, :n] = g.X
        Ri[b, :n, :e] = g.Ri
        Ro[b, :n, :e] = g.Ro
    return X, Ri, Ro
