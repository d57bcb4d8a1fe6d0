#!/usr/bin/env python
"""bench.py -- SegmentClassifier forward throughput (edges/s) on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload acts64|mu200|toy2d]
    python bench.py --impl reference ...      # the reference's CPU algorithm on the host cores

Workloads (BASELINE.json configs):
    acts64  configs[1]: 64 ACTS-like events (~4k hits, ~20k edges each), hidden_dim=32, n_iters=4
    acts64_masked  configs[2]: the same with the masked-linear twin (masks folded in by gnnseg_pack_weights)
    mu200   configs[3]: one mu200-like event (~100k hits, ~1M edges), hidden_dim=64, n_iters=8
    toy2d   configs[0]: 32 Toy2D graphs (40 hits, 144 edges), hidden_dim=8, n_iters=1

A step is one forward pass of the whole batch.  `value` is measured with the batch resident
in HBM (CUDA graph replay, CUDA events per step, L2 flushed between steps).  `e2e` is the
same metric through the public call `model.predict_stream(store.batches(B))` starting from the
HOST event store (gnn_fpga_b200/store.py: the data set narrowed and laid out once at load time,
as the reference loads its graph list once): H2D of the batch's slices, device batch assembly,
forward and D2H of the scores are all inside the timed region; `e2e.from_tuples` is the same from
raw int64 SparseGraph tuples (per-batch host packing inside the timed region too).  The default
single-GPU run also carries a `mu200` sub-record (BASELINE configs[3], the configuration the HBM
target is quoted on).  Multi-GPU: events shard across ranks (weak scaling, every rank its own 64
events), no data-path collective in `value`; `value_with_gather` includes the NCCL all-gather of the
scores, and every rank checks the rows it received against a local recompute; timing is max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "acts64": dict(desc="ACTS-like synthetic events (~4k hits, ~20k edges), hidden_dim=32, n_iters=4, batch 64",
                   F=3, h=32, n_iters=4, batch=64, n_tracks=400, edges_per_hit=5.0),
    "acts64_masked": dict(desc="ACTS-like synthetic events (~4k hits, ~20k edges), hidden_dim=32, n_iters=4, batch 64, "
                               "model_maskedlinear twin (Bernoulli(0.5) masks on the edge / node MLP weights)",
                          F=3, h=32, n_iters=4, batch=64, n_tracks=400, edges_per_hit=5.0, masked=True),
    "mu200": dict(desc="mu200-like synthetic event (~100k hits, ~1M edges), hidden_dim=64, n_iters=8, batch 1",
                  F=3, h=64, n_iters=8, batch=1, n_tracks=10000, edges_per_hit=10.0),
    "toy2d": dict(desc="Toy2D graphs (40 hits, 144 edges), hidden_dim=8, n_iters=1, batch 32",
                  F=3, h=8, n_iters=1, batch=32, n_tracks=4, edges_per_hit=None),
}


def make_graphs(wl, rank):
    from gnn_fpga_b200 import data
    cfg = WORKLOADS[wl]
    if wl == "toy2d":
        return data.toy2d_graphs(cfg["batch"], input_dim=3, seed=rank)
    n_batch = int(os.environ.get("GNNSEG_BENCH_BATCH", cfg["batch"]))     # experiments only: another batch size
    return [data.acts_like_graph(cfg["n_tracks"], seed=rank * cfg["batch"] + b, edges_per_hit=cfg["edges_per_hit"])
            for b in range(n_batch)]


def algorithmic_bytes(Nt, Et, F, h, n_iters):
    """SURVEY.md §8(d) compulsory bytes: every array a pass must touch, once."""
    D = F + h
    inp = Nt * 4 * (F + h)
    edge = Nt * 4 * D + Et * 8 + Et * 4
    node = Nt * 4 * D + Et * 4 + 2 * (Et * 8 + Nt * 4) + Nt * 4 * h
    # hidden_dim = 32 and 64 run the node step as two kernels: the gather touches what §8(d) lists for the node
    # step (its output is the h-wide h1 instead of H'), the MLP reads h1 and X and writes the new state.
    # The forward figure stays §8(d)'s: the h1 round trip is traffic the split adds, not algorithmic bytes.
    return dict(input=inp, edge=edge, node=node, node_gather=node, node_mlp=Nt * 4 * (h + F) + Nt * 4 * D,
                forward=inp + n_iters * (edge + node) + edge)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).
    Started before the warm-up (nvidia-smi takes a moment to come up); only the samples whose
    timestamps fall inside [mark_start, mark_end] are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.t0, self.t1 = [], None, None, None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        samples = []
        for r in self.rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                samples.append((ts, float(r[1]), float(r[2]), r[4:8]))
            except Exception:
                continue
        inside = [x for x in samples if self.t0 is not None and self.t0 - 0.02 <= x[0] <= self.t1 + 0.02]
        if not inside and samples and self.t0 is not None:      # region shorter than the sampling period
            mid = 0.5 * (self.t0 + self.t1)
            inside = [min(samples, key=lambda x: abs(x[0] - mid))]
        reasons = set()
        for x in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), x[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None,
                "sm_max_mhz": inside[0][2] if inside else None, "reasons": sorted(reasons), "samples": len(inside)}


def reference_dense_forward(p, cfg, masks=None):
    """(callable(X, Ri, Ro) -> scores, kind): the reference's own SegmentClassifier (gnn/model.py, placed under
    oracle/_ref/ by oracle/make_ref.py in the build container; kind "reference") when it is there, otherwise the
    oracle's operation-for-operation restatement of it (kind "port", bit-identical on the goldens)."""
    from oracle import segclf_oracle as O
    ref_py = os.path.join(ROOT, "oracle", "_ref", "model.py")
    if os.path.exists(ref_py):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_model", ref_py)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        F, h = cfg["F"], cfg["h"]
        D = F + h
        ones = lambda *shape: torch.ones(*shape)
        me, mn = masks if masks is not None else ([ones(h, 2 * D), ones(1, h)], [ones(h, 3 * D), ones(h, h)])   # gnn/model.py:100 needs masks
        m = mod.SegmentClassifier(input_dim=F, hidden_dim=h, n_iters=cfg["n_iters"], masks_e=me, masks_n=mn)
        m.load_state_dict(p)
        m.eval()

        def fwd(X, Ri, Ro):
            with torch.no_grad():
                return m([X, Ri, Ro])
        return fwd, "reference"
    return (lambda X, Ri, Ro: O.dense_forward(p, X, Ri, Ro, cfg["n_iters"])), "port"


# -----------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own algorithm (dense incidence bmm, gnn/model.py:140-156) restated in
    oracle/segclf_oracle.py, on the host cores.  Rank 0 only.  One step = a bounded sample of
    the workload: ONE event of acts64 (the dense batch does not exist: 41 GB), the full batch
    of toy2d, and for mu200 (800 GB dense) the sparse restatement of the whole event.  Only the
    forward is timed (densifying the event is the generator's job in the reference)."""
    if rank != 0:
        return
    from oracle import segclf_oracle as O
    from gnn_fpga_b200 import graph_from_sparse
    cfg = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    graphs = make_graphs(args.workload, 0)
    p = O.init_params(cfg["F"], cfg["h"], seed=0)
    masks = None
    if cfg.get("masked"):
        from gnn_fpga_b200 import data as _data
        masks = _data.random_masks(cfg["F"], cfg["h"], keep=0.5, seed=1234)
        p = O.apply_masks(p, *masks)
    wl = args.workload
    dense_fwd, kind = reference_dense_forward(p, cfg, masks)

    if wl == "mu200":
        kind = "port"
        X, src, dst, _ = O.flatten_sparse_batch(graphs)
        n_real = int(((src >= 0) & (dst >= 0)).sum())
        sample = "sparse restatement (gather + ordered index_add) of the whole event per step"

        def prepare(i):
            return (X, src, dst), n_real, 1

        def forward(inp):
            O.sparse_forward(p, inp[0], inp[1], inp[2], cfg["n_iters"])
    elif wl in ("acts64", "acts64_masked"):
        what = "gnn/model.py SegmentClassifier.forward itself" if kind == "reference" else "dense restatement of gnn/model.py"
        sample = "%s (incidence bmm), 1 of %d events per step" % (what, len(graphs))

        def prepare(i):
            g = graphs[i % len(graphs)]
            d = graph_from_sparse(g, dtype=np.float32)
            return tuple(torch.from_numpy(a[None]) for a in (d.X, d.Ri, d.Ro)), g.Ri_rows.shape[0], 1

        def forward(inp):
            dense_fwd(inp[0], inp[1], inp[2])
    else:
        what = "gnn/model.py SegmentClassifier.forward itself" if kind == "reference" else "dense restatement of gnn/model.py"
        sample = "%s (incidence bmm), the full batch of %d graphs per step" % (what, len(graphs))
        dense = tuple(torch.from_numpy(a) for a in O.merge_dense([graph_from_sparse(g) for g in graphs]))
        n_real = sum(g.Ri_rows.shape[0] for g in graphs)

        def prepare(i):
            return dense, n_real, len(graphs)

        def forward(inp):
            dense_fwd(inp[0], inp[1], inp[2])

    for i in range(args.warmup):
        forward(prepare(i)[0])
    edges, events, dt = 0, 0, 0.0
    for i in range(args.steps):
        inp, n_e, n_ev = prepare(args.warmup + i)
        t0 = time.perf_counter()
        forward(inp)
        dt += time.perf_counter() - t0
        edges += n_e
        events += n_ev
    value = edges / dt
    line = {
        "impl": "reference", "metric": "segment_classifier_forward_edges_per_sec", "value": value, "unit": "edges/s",
        "events_per_sec": events / dt, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "hidden_dim": cfg["h"], "n_iters": cfg["n_iters"], "batch": cfg["batch"]},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(workload, budget_s=15.0):
    """Bounded CPU sample of the same workload with the oracle, on this box's host cores."""
    from oracle import segclf_oracle as O
    from gnn_fpga_b200 import graph_from_sparse
    cfg = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    graphs = make_graphs(workload, 0)
    p = O.init_params(cfg["F"], cfg["h"], seed=0)
    masks = None
    if cfg.get("masked"):
        from gnn_fpga_b200 import data as _data
        masks = _data.random_masks(cfg["F"], cfg["h"], keep=0.5, seed=1234)
        p = O.apply_masks(p, *masks)
    dense_fwd, kind = reference_dense_forward(p, cfg, masks)
    out = {"unit": "edges/s", "cores": cores, "kind": "port" if workload == "mu200" else kind}
    X, src, dst, _ = O.flatten_sparse_batch(graphs)
    n_real = int(((src >= 0) & (dst >= 0)).sum())
    t0 = time.perf_counter()
    O.sparse_forward(p, X, src, dst, cfg["n_iters"])
    ts = time.perf_counter() - t0
    out["sparse_port"] = {"value": n_real / ts, "unit": "edges/s", "sample": "whole batch once, sparse restatement"}
    if workload == "mu200":
        out["value"] = n_real / ts
        out["sample"] = "sparse restatement, whole event once (the dense reference needs 800 GB)"
        return out
    edges, n, t_total = 0, 0, 0.0
    per_event = workload in ("acts64", "acts64_masked")
    while t_total < budget_s and n < (len(graphs) if per_event else 50):
        if per_event:
            g = graph_from_sparse(graphs[n], dtype=np.float32)
            Xd, Ri, Ro = (torch.from_numpy(a[None]) for a in (g.X, g.Ri, g.Ro))
            e = graphs[n].Ri_rows.shape[0]
        else:
            Xd, Ri, Ro = (torch.from_numpy(a) for a in O.merge_dense([graph_from_sparse(g) for g in graphs]))
            e = n_real
        t0 = time.perf_counter()
        dense_fwd(Xd, Ri, Ro)
        t_total += time.perf_counter() - t0
        edges += e
        n += 1
    out["value"] = edges / t_total
    out["sample"] = ("%s, %d %s in %.1f s" % ("gnn/model.py SegmentClassifier.forward itself" if kind == "reference" else
                                              "dense restatement of gnn/model.py", n,
                                              "events one at a time" if per_event else "full batches", t_total))
    return out


def mu200_throughput(dev, steps, warmup, n_events=4):
    """configs[3] asks for single-event latency AND throughput: the same forward on a batch of `n_events` mu200-like
    events (batch resident, CUDA graph, L2 flushed between steps), so that the fixed cost per launch is shared."""
    from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, GraphStore, data
    cfg = WORKLOADS["mu200"]
    graphs = [data.acts_like_graph(cfg["n_tracks"], seed=b, edges_per_hit=cfg["edges_per_hit"]) for b in range(n_events)]
    torch.manual_seed(0)
    model = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"]).to(dev).eval()
    batch = DeviceGraphBatch.from_store(GraphStore.from_sparse_graphs(graphs, reorder="none"), 0, n_events, dev)
    n_real = batch.count_real_edges()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps = min(steps, 50)
    with torch.no_grad():
        for _ in range(warmup):
            model(batch)
        torch.cuda.synchronize(dev)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            model._run(batch)
            ends[i].record()
        torch.cuda.synchronize(dev)
        model.check_range()
    ms = float(sum(s_.elapsed_time(e_) for s_, e_ in zip(starts, ends))) / steps
    return {"batch": n_events, "ms_per_step": ms, "ms_per_event": ms / n_events, "value": n_real / (ms * 1e-3), "unit": "edges/s",
            "events_per_sec": n_events / (ms * 1e-3), "steps": steps}


# -----------------------------------------------------------------------------------------
def measure_workload(workload, args, rank, world, local_rank, dev, dist, sampler=None, light=False):
    """Everything bench.py reports for one workload on this rank: device-timed forward (`value`), per-kernel
    times, e2e through predict_stream from a GraphStore (and from raw tuples), the training step.  Returns a
    dict of per-rank numbers (reduced over ranks by the caller).  light = True: the sub-record of a second
    workload (no training step, fewer extras)."""
    from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, GraphStore, _lib
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    import ctypes as C

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    cfg = WORKLOADS[workload]
    graphs = make_graphs(workload, rank)
    torch.manual_seed(0)
    masks_e = masks_n = None
    if cfg.get("masked"):
        from gnn_fpga_b200 import data as _data
        masks_e, masks_n = _data.random_masks(cfg["F"], cfg["h"], keep=0.5, seed=1234)
    model = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"], masks_e=masks_e, masks_n=masks_n).to(dev).eval()
    # load time: the data set goes into the pinned event store once (narrowed, validated, laid out)
    t0 = time.perf_counter()
    store = GraphStore.from_sparse_graphs(graphs, reorder=os.environ.get("GNNSEG_BENCH_REORDER", "none"))
    store_build_s = time.perf_counter() - t0
    n_events = len(graphs)
    batch = DeviceGraphBatch.from_store(store, 0, n_events, dev)
    n_real = batch.count_real_edges()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    steps = args.steps
    h, F, it = cfg["h"], cfg["F"], cfg["n_iters"]
    fused = h in (32, 64) and not os.environ.get("GNNSEG_EXACT")

    # ---- value: batch resident in HBM ----------------------------------------------------
    with torch.no_grad():
        for _ in range(args.warmup):
            model(batch)
        barrier()
        if sampler:
            sampler.mark_start()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            model._run(batch)                 # the CUDA-graph replay alone (model(batch) adds a clone of the scores)
            ends[i].record()
        barrier()
        if sampler:
            sampler.mark_end()
        model.check_range()
    step_ms = [s_.elapsed_time(e_) for s_, e_ in zip(starts, ends)]
    res = {"total_ms": float(sum(step_ms)), "n_real": n_real, "n_events": n_events, "n_nodes": batch.n_nodes,
           "store_build_ms": store_build_s * 1e3, "fused": fused}
    # pack x2 + status memset, input, then per iteration: fused path 2 kernels; step-by-step path edge + node (2 kernels at h = 32 / 64)
    if fused:
        res["launches_per_step"] = 2 + 1 + 2 * it + 1
    else:
        res["launches_per_step"] = 2 + 1 + (it + 1) + it * (2 if h in (32, 64) else 1)

    # ---- per-kernel durations (same process, CUDA events around single launches) ----------
    L = _lib.lib()
    blob = model.pack_weights()
    n, m = batch.n_nodes, batch.n_slots
    X4 = torch.empty(n, 4, device=dev)
    kt = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rc = fn(); b.record()
        assert rc == 0, (name, rc)
        kt.setdefault(name, []).append((a, b))

    st = _stream_ptr(dev)
    reps = max(3, min(steps, 20))
    gs = C.byref(batch.struct)
    if fused:
        S = [torch.zeros(n + 1, 5 * h, device=dev) for _ in range(2)]       # row n: what an absent neighbour reads
        sc = torch.empty(m, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        for rep in range(reps + 1):
            if rep == 1:
                kt.clear()
            S[0][n].zero_(); S[1][n].zero_()
            flush.zero_()
            timed("input", lambda: L.gnnseg_state_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(S[0]),
                                                             5 * h if it else 2 * h, _ptr(status), st))
            cur = 0
            for i in range(it):
                last = i + 1 == it
                rows = S[cur ^ 1]
                timed("fused_gather", lambda: L.gnnseg_fused_gather_step(_ptr(blob), gs, _ptr(S[cur]), h, _ptr(rows), 5 * h, st))
                timed("node_mlp", lambda: L.gnnseg_state_mlp_step(_ptr(blob), _ptr(X4), _ptr(rows), 5 * h, n, h, _ptr(rows),
                                                                  2 * h if last else 5 * h, _ptr(status), st))
                cur ^= 1
            S[cur][n, :h] = blob[-h:]                       # 2^(log2e b1), the last h floats of the packed weights
            timed("edge_final", lambda: L.gnnseg_edge_final_step(_ptr(blob), gs, _ptr(S[cur]), 5 * h, 0, h, h, _ptr(sc), st))
        assert torch.equal(sc, batch.scores), "per-kernel replay of the forward differs from gnnseg_forward_ex"
        mult = {"input": 1, "edge_final": 1}
    else:
        Q = [torch.empty(n, 3 * h, device=dev) for _ in range(2)]
        P = torch.empty(n, 2 * h, device=dev)
        e = torch.empty(m, device=dev)
        e_in = torch.empty(m, device=dev)
        e_out = torch.empty(m, device=dev)
        split = h in (32, 64)
        h1 = torch.empty(n, h, device=dev) if split else None
        for rep in range(reps + 1):
            if rep == 1:
                kt.clear()
            flush.zero_()
            timed("input", lambda: L.gnnseg_input_step(_ptr(blob), _ptr(batch.X), n, F, h, _ptr(X4), _ptr(P), _ptr(Q[0]), st))
            cur = 0
            for i in range(it):
                qo = _ptr(Q[cur ^ 1]) if i + 1 < it else None
                timed("edge", lambda: L.gnnseg_edge_step(_ptr(blob), gs, _ptr(P), h, None, _ptr(e_in), _ptr(e_out), st))
                if split:
                    timed("node_gather", lambda: L.gnnseg_node_gather_step(gs, _ptr(Q[cur]), _ptr(e_in), _ptr(e_out), h, _ptr(h1), h, st))
                    timed("node_mlp", lambda: L.gnnseg_node_mlp_step(_ptr(blob), _ptr(X4), _ptr(h1), h, n, h, _ptr(P), qo, st))
                else:
                    timed("node", lambda: L.gnnseg_node_step(_ptr(blob), gs, _ptr(X4), _ptr(Q[cur]), _ptr(e_in), _ptr(e_out), h, _ptr(P), qo, st))
                cur ^= 1
            timed("edge", lambda: L.gnnseg_edge_step(_ptr(blob), gs, _ptr(P), h, _ptr(e), None, None, st))
        mult = {"input": 1, "edge": it + 1}
    torch.cuda.synchronize(dev)
    res["kernel_ms"] = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in kt.items() if v}
    res["kernel_share"] = {k: res["kernel_ms"][k] * mult.get(k, it) for k in res["kernel_ms"]}

    # ---- e2e: host -> scores on the host, through model.predict_stream ----------------------------
    if not args.no_e2e:
        sb = store.batch(0, n_events)
        host_ref = None
        with torch.no_grad():
            for out_host in model.predict_stream([sb] * 6):      # every pipeline slot allocated and warm (kept on the model afterwards)
                host_ref = out_host.clone()
            barrier()
            t0 = time.perf_counter()
            n_out = 0
            for out_host in model.predict_stream([sb] * steps):
                n_out += 1
            torch.cuda.synchronize(dev)
            sec_store = time.perf_counter() - t0
            assert n_out == steps and torch.equal(out_host, host_ref)
            assert torch.equal(host_ref, batch.scores.view(n_events, batch.e_max).cpu())      # same bits as the resident run
            # stage breakdown of one batch, serialised (CUDA events between the stages)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            pinned_out = torch.empty((n_events, batch.e_max), dtype=torch.float32, pin_memory=True)
            stages = np.zeros(3)
            host_enqueue = 0.0
            cur_stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            bufs = batch._bufs
            for rep in range(4):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                ev[0].record()
                assert L.gnnseg_store_load_batch(C.byref(store.layout), store.arena.data_ptr(), 0, n_events, C.byref(bufs.struct()),
                                                 cur_stream, cur_stream, None) == 0          # H2D + assembly + adjacency
                ev[1].record()
                sc2 = model._run(batch)
                ev[2].record()
                pinned_out.copy_(sc2.view(n_events, batch.e_max), non_blocking=True)
                ev[3].record()
                host_enqueue = time.perf_counter() - t0
                torch.cuda.synchronize(dev)
                if rep > 0:
                    stages += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
            stages /= 3
            # host time of one pipelined batch: one library call (copies, assembly, forward, D2H enqueued) + two events
            t0 = time.perf_counter()
            n_rep = 20
            for out_host in model.predict_stream([sb] * n_rep):
                pass
            per_batch_wall = (time.perf_counter() - t0) / n_rep
            res["e2e"] = {"sec": sec_store, "h2d": store.h2d_bytes(0, n_events), "d2h": pinned_out.numel() * 4 + 4,
                          "stages_ms": {"host_pack": 0.0, "h2d_and_assemble": float(stages[0]), "forward": float(stages[1]),
                                        "d2h": float(stages[2]), "host_enqueue_unpipelined": host_enqueue * 1e3,
                                        "pipelined_wall_per_batch": per_batch_wall * 1e3}}
            if not light:
                # the same from raw int64 SparseGraph tuples: a worker thread narrows / packs every batch first
                for _ in model.predict_stream([graphs] * 2):
                    pass
                barrier()
                t0 = time.perf_counter()
                for out_host in model.predict_stream([graphs] * max(3, steps // 4)):
                    pass
                torch.cuda.synchronize(dev)
                res["e2e"]["sec_tuples"] = (time.perf_counter() - t0) / max(3, steps // 4) * steps
                assert torch.equal(out_host, host_ref)
                t0 = time.perf_counter()
                for _ in range(max(3, steps // 4)):
                    pinned_out.copy_(model(graphs), non_blocking=True)
                    torch.cuda.synchronize(dev)
                res["e2e"]["sec_blocking"] = (time.perf_counter() - t0) / max(3, steps // 4) * steps

    # ---- training step (BASELINE configs[4]): forward_train + BCE + L1 + backward + all-reduce + Adam ----
    if not args.no_train and not light:
        from gnn_fpga_b200.training import NativeTrainer
        tmodel = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"], masks_e=masks_e, masks_n=masks_n).to(dev).train()
        tmodel.load_state_dict(model.state_dict())
        y = store.targets(0, n_events).to(dev)
        trainer = NativeTrainer(tmodel, l1=1e-4)
        for _ in range(3):
            trainer.step(batch, y)
        barrier()
        n_tr = max(3, min(steps, 20))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_tr):
            trainer.step(batch, y)
        b.record()
        barrier()
        res["train"] = {"ms": a.elapsed_time(b) / n_tr, "loss": float(trainer.loss.item())}
        del trainer, tmodel
        batch._ws = {k: v for k, v in batch._ws.items() if not (isinstance(k, tuple) and k[0] == "train")}

    # ---- the collective of the inference path: all ranks receive all scores (NCCL all-gather) --------
    if world > 1 and not light:
        from gnn_fpga_b200.dist import ScoreGatherer
        gat = ScoreGatherer()
        local = batch.scores_2d()
        with torch.no_grad():
            allsc = gat(local)                                  # exchanges the block shapes once
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                allsc = gat(local)
            b.record()
            torch.cuda.synchronize(dev)
            res["gather_ms"] = a.elapsed_time(b) / 10
            # value_with_gather: forward + all-gather back to back on the compute stream, L2 flushed per step
            barrier()
            sg = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            eg = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            for i in range(steps):
                flush.zero_()
                sg[i].record()
                model._run(batch)
                allsc = gat(batch.scores_2d())
                eg[i].record()
            barrier()
            res["total_ms_with_gather"] = float(sum(s_.elapsed_time(e_) for s_, e_ in zip(sg, eg)))
            # SURVEY 8(e) gate: the rows rank r receives from rank (r + 1) % world are bit-equal to this rank's own
            # forward of that rank's events (deterministic kernels, independent events)
            other = (rank + 1) % world
            ob = DeviceGraphBatch.from_store(GraphStore.from_sparse_graphs(make_graphs(workload, other), reorder=os.environ.get("GNNSEG_BENCH_REORDER", "none")),
                                             0, n_events, dev)
            mine = model._run(ob).view(n_events, ob.e_max)
            got = allsc[other * n_events:(other + 1) * n_events, :ob.e_max]
            res["gather_bit_equal"] = bool(torch.equal(mine, got))
            assert res["gather_bit_equal"], "rank %d: gathered scores of rank %d differ from a local recompute" % (rank, other)
    res["cfg"] = cfg
    return res


def reduce_over_ranks(res, world, dist, dev):
    """max over ranks of the times, sum over ranks of the counts (in place)."""
    if world <= 1:
        res["all_edges"], res["all_events"] = float(res["n_real"]), float(res["n_events"])
        return res
    e2e = res.get("e2e", {})
    times = torch.tensor([res["total_ms"], e2e.get("sec", 0.0), e2e.get("sec_tuples", 0.0), e2e.get("sec_blocking", 0.0),
                          res.get("gather_ms", 0.0), res.get("total_ms_with_gather", 0.0), res.get("train", {}).get("ms", 0.0)],
                         dtype=torch.float64, device=dev)
    counts = torch.tensor([res["n_real"], res["n_events"]], dtype=torch.float64, device=dev)
    dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    t = times.tolist()
    res["total_ms"] = t[0]
    if e2e:
        e2e["sec"], e2e["sec_tuples"], e2e["sec_blocking"] = t[1], t[2], t[3]
    if "gather_ms" in res:
        res["gather_ms"], res["total_ms_with_gather"] = t[4], t[5]
    if "train" in res:
        res["train"]["ms"] = t[6]
    res["all_edges"], res["all_events"] = counts.tolist()
    return res


def record_of(res, args, world, workload, peak, peak_src):
    """The JSON fields of one workload from its (rank-reduced) measurements."""
    cfg = res["cfg"]
    h, F, it = cfg["h"], cfg["F"], cfg["n_iters"]
    steps = args.steps
    ab = algorithmic_bytes(res["n_nodes"], res["n_real"], F, h, it)
    ab["fused_gather"] = ab["edge"] + ab["node"]        # the edge step evaluated inside the node step's CSR walk
    ab["edge_final"] = ab["edge"]
    kernel_ms, share = res["kernel_ms"], res["kernel_share"]
    dom = max(share, key=share.get)
    try:     # DRAM bytes per launch of that kernel, from the committed ncu capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[workload][dom]
    except Exception:
        traffic = None
    achieved = ab[dom] / (kernel_ms[dom] * 1e-3) / 1e9
    ms_per_step = res["total_ms"] / steps
    fwd_gbs = ab["forward"] / (ms_per_step * 1e-3) / 1e9
    rec = {
        "value": res["all_edges"] * steps / (res["total_ms"] * 1e-3), "unit": "edges/s",
        "events_per_sec": res["all_events"] * steps / (res["total_ms"] * 1e-3), "ms_per_step": ms_per_step,
        "config": {"workload": cfg["desc"], "hidden_dim": h, "n_iters": it, "batch_per_gpu": res["n_events"],
                   "nodes_per_gpu": res["n_nodes"], "edges_per_gpu": res["n_real"], "parallelism": "events sharded, dp%d" % world,
                   "l2": "flushed between timed steps (256 MiB write)", "cuda_graph": True,
                   "path": "fused (edge step inside the node step's CSR walk)" if res["fused"] else "step by step"},
        "roofline": {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ab[dom], "kernel_ms": kernel_ms[dom]},
        "roofline_forward": {"achieved": fwd_gbs, "frac": fwd_gbs / peak, "unit": "GB/s", "algorithmic_bytes": ab["forward"]},
        "kernel_ms": kernel_ms, "kernel_ms_per_step": share,
        "gpu_launches": res["launches_per_step"] * steps,
    }
    if "e2e" in res:
        e = res["e2e"]
        rec["e2e"] = {"value": res["all_edges"] * steps / e["sec"], "unit": "edges/s",
                      "h2d_bytes_per_step": e["h2d"], "d2h_bytes_per_step": e["d2h"], "ms_per_step": e["sec"] / steps * 1e3,
                      "path": "model.predict_stream(store.batches(B)): the data set sits in a GraphStore (pinned host arena, built once at "
                              "load time: %.1f ms for this batch); per batch ONE library call (gnnseg_store_forward_batch): six H2D copies "
                              "on a copy stream followed there by the (lean) batch assembly, the forward on the compute stream, D2H of the scores into pinned memory on a "
                              "third stream; three batches in flight" % res["store_build_ms"],
                      "stages_ms": e["stages_ms"]}
        if e.get("sec_tuples"):
            rec["e2e"]["from_tuples"] = {"value": res["all_edges"] * steps / e["sec_tuples"], "ms_per_step": e["sec_tuples"] / steps * 1e3,
                                         "path": "model.predict_stream(batches of raw int64 SparseGraph tuples): a worker thread narrows and "
                                                 "validates every batch into a one-batch store first (<= 4 threads), then the same device path"}
            rec["e2e"]["blocking"] = {"value": res["all_edges"] * steps / e["sec_blocking"], "ms_per_step": e["sec_blocking"] / steps * 1e3,
                                      "path": "model(graphs) then copy to pinned host memory, synchronised every step"}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="acts64", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-mu200", action="store_true", help="skip the mu200 sub-record of the default single-GPU run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpus = None
    if world > 1:
        if os.environ.get("GNNSEG_BIND_CPUS", "1") != "0":       # pinned host memory on the GPU's own NUMA node
            from gnn_fpga_b200.dist import bind_to_local_cpus
            cpus = bind_to_local_cpus(local_rank)
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = measure_workload(args.workload, args, rank, world, local_rank, dev, dist, sampler=sampler)
    clocks = sampler.stop() if sampler else None
    reduce_over_ranks(res, world, dist, dev)
    # the configuration the HBM target is quoted on (BASELINE configs[3]) rides along in the default single-GPU run
    sub = None
    if world == 1 and args.workload == "acts64" and not args.no_mu200:
        sub = reduce_over_ranks(measure_workload("mu200", args, rank, world, local_rank, dev, dist, light=True), world, dist, dev)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        rec = record_of(res, args, world, args.workload, peak, peak_src)
        cfg = res["cfg"]
        it, h = cfg["n_iters"], cfg["h"]
        line = {"metric": "segment_classifier_forward_edges_per_sec", "value": rec.pop("value"), "unit": rec.pop("unit"),
                "events_per_sec": rec.pop("events_per_sec"), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic"}
        line.update(rec)
        line["clocks"] = clocks
        if "train" in res:
            tr = res["train"]
            # backward: edge, gather and the dense step per (iteration + final step); the dense step is two kernels (tcgen05) at hidden_dim 32 / 64
            n_k = 2 + 1 + (it + 1) + it * (2 if h in (32, 64) else 1) + 2 + 1 + (4 if h in (32, 64) else 3) * (it + 1) + 1 + 1 + 1
            line["train_step"] = {"ms": tr["ms"], "edges_per_sec": res["all_edges"] / (tr["ms"] * 1e-3),
                                  "events_per_sec": res["all_events"] / (tr["ms"] * 1e-3), "loss": tr["loss"],
                                  "gpu_launches_per_step": n_k,
                                  "path": "NativeTrainer.step on the resident batch: gnnseg_forward_train, gnnseg_bce_loss, "
                                          "gnnseg_backward (dense step on tcgen05 at hidden_dim 32 / 64), gnnseg_l1_penalty, %sgnnseg_adam_step; L2 not flushed"
                                          % ("NCCL all-reduce of the flat gradient, " if world > 1 else "")}
        if world > 1:
            line["scores_allgather_ms"] = res["gather_ms"]      # one all_gather_into_tensor of every rank's (B, E_max) scores
            line["value_with_gather"] = res["all_edges"] * args.steps / (res["total_ms_with_gather"] * 1e-3)
            line["ms_per_step_with_gather"] = res["total_ms_with_gather"] / args.steps
            line["gathered_scores_bit_equal_to_local_recompute"] = res.get("gather_bit_equal")
            line["rank0_cpu_binding"] = ("%d cores next to the GPU (gnn_fpga_b200.dist.bind_to_local_cpus)" % len(cpus)) if cpus else "none"
        if sub is not None:
            line["mu200"] = record_of(sub, args, world, "mu200", peak, peak_src)
            line["mu200"]["throughput_batch4"] = mu200_throughput(dev, args.steps, args.warmup)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
            if sub is not None:
                line["mu200"]["cpu_baseline"] = cpu_baseline("mu200")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
