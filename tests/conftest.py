import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = [
    "c1_toy2d_h8_it1", "toy2d_f2_h32_it10", "toy2d_h4_it1", "toy2d_h16_it3", "acts_ragged_h32_it4",
    "acts_h64_it6", "acts_masked_h8_it4", "acts_masked_h8_it4_twin", "acts_masked_h32_it4",
    "half_edges_h8_it2",
]


NODECLF_CASES = ["nodeclf_toy2d_h8_it1", "nodeclf_toy2d_h8_it0", "nodeclf_f4_h16_it2", "nodeclf_acts_h32_it3",
                 "nodeclf_acts_h64_it2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (*.so is git-ignored): build it once (nvcc cross-compiles
    without a GPU).  Nothing happens when the library is already there."""
    if not os.path.exists(os.path.join(ROOT, "gnn_fpga_b200", "libgnnseg_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


def load_case(name):
    """Golden record written by oracle/make_golden.py (outputs of the reference itself)."""
    import torch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    rec["params"] = {k[len("param:"):]: torch.from_numpy(v.copy()) for k, v in rec.items() if k.startswith("param:")}
    rec["grads"] = {k[len("grad:"):]: v for k, v in rec.items() if k.startswith("grad:")}
    for k in ("F", "h", "n_iters", "seed", "n_params"):
        rec[k] = int(rec[k])
    if "mask_e0" in rec:
        rec["masks_e"] = [torch.from_numpy(rec["mask_e0"].copy()), torch.from_numpy(rec["mask_e1"].copy())]
        rec["masks_n"] = [torch.from_numpy(rec["mask_n0"].copy()), torch.from_numpy(rec["mask_n1"].copy())]
    else:
        rec["masks_e"] = rec["masks_n"] = None
    return rec


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
