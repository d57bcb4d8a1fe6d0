#!/usr/bin/env python
"""What does node locality buy the gathers?  Relabels the nodes of every synthetic event on the HOST
(edges and scores keep their slots) and runs bench.py's device-timed part on the result.

    python scripts/locality_experiment.py [acts64|mu200] [none|phi|layer_phi|random] ...

order = none       the generator's order (layer major, tracks in random phi order inside a layer)
        phi        nodes sorted by phi alone
        layer_phi  nodes sorted by (layer, phi)
        random     a random permutation (worst case)
One JSON line per variant on stdout (bench.py's line, cpu / e2e / training legs off).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402
from gnn_fpga_b200.data import sparse_from_edges  # noqa: E402


def relabel(g, order):
    if order == "none":
        return g
    n = g.X.shape[0]
    if order == "phi":
        key = g.X[:, 1].astype(np.float64)
    elif order == "layer_phi":
        key = np.round(g.X[:, 0].astype(np.float64) * 1000.0) * 10.0 + g.X[:, 1].astype(np.float64)
    else:
        key = np.random.RandomState(n).rand(n)
    perm = np.argsort(key, kind="stable")       # new position -> old node
    rank = np.empty(n, np.int64)
    rank[perm] = np.arange(n)                    # old node -> new position
    src = np.empty(g.Ro_cols.shape[0], np.int64); src[g.Ro_cols] = g.Ro_rows
    dst = np.empty(g.Ri_cols.shape[0], np.int64); dst[g.Ri_cols] = g.Ri_rows
    return sparse_from_edges(g.X[perm], rank[src], rank[dst], g.y)


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "acts64"
    orders = sys.argv[2:] or ["none", "phi", "layer_phi", "random"]
    base_make = bench.make_graphs
    for order in orders:
        bench.make_graphs = lambda wl, rank, _o=order: [relabel(g, _o) for g in base_make(wl, rank)]
        sys.argv = ["bench.py", "--workload", workload, "--steps", "30", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--no-train"]
        print("## order=%s" % order, flush=True)
        bench.main()


if __name__ == "__main__":
    main()
