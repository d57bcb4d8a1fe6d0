"""Full-size parity on the configurations the targets are quoted on (BASELINE.json configs[1], [3])
and the large-weight error report SURVEY.md §7.3 asks for.  Everything here needs a GPU; the CPU
side (the oracle's sparse restatement, fp32 and fp64) takes a few seconds per case.

Reference: gnn/model.py:140-156 (forward), config gnn/MPNN_Seg_ACTS_mu200.ipynb:282-286 (h=64).
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import segclf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _model(p, F, h, n_iters, device):
    from gnn_fpga_b200 import SegmentClassifier
    m = SegmentClassifier(F, h, n_iters)
    m.load_state_dict(p)
    return m.to(device).eval()


def _report(name, rec):
    """Append one record to gpurun_out/parity_report.jsonl when the scratch directory exists."""
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(case=name, **rec)) + "\n")
    print(name, rec)


def test_mu200_full_size_against_sparse_oracle(cuda_device):
    """BASELINE configs[3] at full size: one mu200-like event (100 k hits, ~1 M edges), hidden_dim=64,
    n_iters=8.  Scores vs the oracle's sparse restatement in fp32 and fp64: <= 1e-5 relative."""
    from gnn_fpga_b200 import data
    g = data.mu200_like_graph(seed=0)
    assert g.X.shape[0] == 100000 and g.Ri_rows.shape[0] > 900000
    p = O.init_params(3, 64, seed=0)
    model = _model(p, 3, 64, 8, cuda_device)
    with torch.no_grad():
        out = model([g])[0].cpu().numpy()
    X, src, dst, e_max = O.flatten_sparse_batch([g])
    ref32 = O.sparse_forward(p, X, src, dst, 8, torch.float32).numpy()
    ref64 = O.sparse_forward(p, X, src, dst, 8, torch.float64).numpy()
    e32, e64 = rel_err(out, ref32), rel_err(out, ref64)
    _report("mu200_full_h64_it8", {"edges": int(e_max), "rel_err_vs_fp32_oracle": e32, "rel_err_vs_fp64_oracle": e64,
                                   "fp32_oracle_vs_fp64": rel_err(ref32, ref64)})
    assert e32 <= TOL and e64 <= TOL


def test_acts64_batch_events_against_oracle(cuda_device):
    """BASELINE configs[1] at full size (64 events in one batch, ~1.29 M edges, hidden_dim=32, n_iters=4):
    three events taken out of the batch vs the oracle run on each event alone (events are independent)."""
    from gnn_fpga_b200 import data
    graphs = data.acts_like_graphs(64, 400, seed=0)
    p = O.init_params(3, 32, seed=0)
    model = _model(p, 3, 32, 4, cuda_device)
    with torch.no_grad():
        out = model(graphs).cpu().numpy()
    worst = 0.0
    for b in (0, 31, 63):
        X, src, dst, e_max = O.flatten_sparse_batch([graphs[b]])
        ref32 = O.sparse_forward(p, X, src, dst, 4, torch.float32).numpy()
        ref64 = O.sparse_forward(p, X, src, dst, 4, torch.float64).numpy()
        worst = max(worst, rel_err(out[b, :e_max], ref32), rel_err(out[b, :e_max], ref64))
    _report("acts64_batch_events_0_31_63", {"rel_err": worst})
    assert worst <= TOL


@pytest.mark.parametrize("h,n_iters,scale", [(32, 4, 4.0), (64, 6, 4.0), (8, 4, 8.0)])
def test_large_weights_error_report(h, n_iters, scale, cuda_device):
    """Trained-size weights (|w| up to 15 in the reference's trained models): every
    weight matrix scaled by `scale`, scores saturate towards 0 and 1.  Error against the fp64
    restatement for (a) the CUDA path and (b) the reference's own fp32 algorithm (dense restatement,
    bit-exact with gnn/model.py on the goldens).  Both are fp32 evaluations of a saturated, ill-conditioned
    network, so each is one sample of the rounding noise: SURVEY.md 7.3 measured the reference 4-9e-6 from the fp64
    truth and two fp32 orders up to 1.3e-5 apart, i.e. a factor of 3 between samples; on top of that the tensor-core
    GEMMs (hidden_dim 32 / 64) are 3xTF32, which drops the lo x lo term (2^-22 relative per product against fp32's
    2^-24).  The gate: the worst edge and the 99th percentile within 6x of the reference's own (floor 2e-6); both
    figures and the largest absolute error go into the parity report (measured: 2.4x - 4.8x)."""
    from gnn_fpga_b200 import data, graph_from_sparse
    worst_cuda = worst_ref = worst_abs = 0.0
    p99_cuda = p99_ref = 0.0
    for seed in (0, 1, 2):
        g = data.acts_like_graph(40, seed=seed)
        p = {k: (v * scale if k.endswith("weight") else v) for k, v in O.init_params(3, h, seed=seed).items()}
        model = _model(p, 3, h, n_iters, cuda_device)
        with torch.no_grad():
            out = model([g])[0].cpu().numpy()
        X, src, dst, e_max = O.flatten_sparse_batch([g])
        ref64 = O.sparse_forward(p, X, src, dst, n_iters, torch.float64).numpy()
        d = graph_from_sparse(g, dtype=np.float32)
        dense32 = O.dense_forward(p, *(torch.from_numpy(a[None]) for a in (d.X, d.Ri, d.Ro)), n_iters)[0].numpy()
        assert np.all(np.isfinite(out))
        worst_cuda = max(worst_cuda, rel_err(out, ref64))
        worst_ref = max(worst_ref, rel_err(dense32, ref64))
        worst_abs = max(worst_abs, float(np.max(np.abs(out.astype(np.float64) - ref64))))
        p99_cuda = max(p99_cuda, float(np.percentile(np.abs(out.astype(np.float64) - ref64) / np.abs(ref64), 99)))
        p99_ref = max(p99_ref, float(np.percentile(np.abs(dense32.astype(np.float64) - ref64) / np.abs(ref64), 99)))
        assert float(ref64.min()) < 0.2 and float(ref64.max()) > 0.8      # the regime the test is about
    _report("large_weights_h%d_it%d_x%g" % (h, n_iters, scale),
            {"cuda_rel_err_vs_fp64": worst_cuda, "reference_fp32_rel_err_vs_fp64": worst_ref, "cuda_abs_err_vs_fp64": worst_abs,
             "cuda_p99_rel_err": p99_cuda, "reference_fp32_p99_rel_err": p99_ref})
    assert worst_cuda <= 6.0 * max(worst_ref, 2e-6)
    assert p99_cuda <= 6.0 * max(p99_ref, 2e-6)
