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
same metric through the public call `model(list_of_SparseGraph)` starting from HOST numpy
tuples: host packing, H2D, device CSR build, forward and D2H of the scores are all inside the
timed region.  Multi-GPU: events shard across ranks (weak scaling, every rank its own 64
events), no data-path collective; timing is max over ranks.
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
    return [data.acts_like_graph(cfg["n_tracks"], seed=rank * cfg["batch"] + b, edges_per_hit=cfg["edges_per_hit"])
            for b in range(cfg["batch"])]


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
    if cfg.get("masked"):
        from gnn_fpga_b200 import data as _data
        p = O.apply_masks(p, *_data.random_masks(cfg["F"], cfg["h"], keep=0.5, seed=1234))
    wl = args.workload

    if wl == "mu200":
        X, src, dst, _ = O.flatten_sparse_batch(graphs)
        n_real = int(((src >= 0) & (dst >= 0)).sum())
        sample = "sparse restatement (gather + ordered index_add) of the whole event per step"

        def prepare(i):
            return (X, src, dst), n_real, 1

        def forward(inp):
            O.sparse_forward(p, inp[0], inp[1], inp[2], cfg["n_iters"])
    elif wl in ("acts64", "acts64_masked"):
        sample = "dense restatement of gnn/model.py (incidence bmm), 1 of %d events per step" % len(graphs)

        def prepare(i):
            g = graphs[i % len(graphs)]
            d = graph_from_sparse(g, dtype=np.float32)
            return tuple(torch.from_numpy(a[None]) for a in (d.X, d.Ri, d.Ro)), g.Ri_rows.shape[0], 1

        def forward(inp):
            O.dense_forward(p, inp[0], inp[1], inp[2], cfg["n_iters"])
    else:
        sample = "dense restatement of gnn/model.py (incidence bmm), the full batch of %d graphs per step" % len(graphs)
        dense = tuple(torch.from_numpy(a) for a in O.merge_dense([graph_from_sparse(g) for g in graphs]))
        n_real = sum(g.Ri_rows.shape[0] for g in graphs)

        def prepare(i):
            return dense, n_real, len(graphs)

        def forward(inp):
            O.dense_forward(p, inp[0], inp[1], inp[2], cfg["n_iters"])

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
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": "port", "sample": sample},
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
    if cfg.get("masked"):
        from gnn_fpga_b200 import data as _data
        p = O.apply_masks(p, *_data.random_masks(cfg["F"], cfg["h"], keep=0.5, seed=1234))
    out = {"unit": "edges/s", "cores": cores, "kind": "port"}
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
        O.dense_forward(p, Xd, Ri, Ro, cfg["n_iters"])
        t_total += time.perf_counter() - t0
        edges += e
        n += 1
    out["value"] = edges / t_total
    out["sample"] = ("dense restatement of gnn/model.py, %d %s in %.1f s" %
                     (n, "events one at a time" if per_event else "full batches", t_total))
    return out


# -----------------------------------------------------------------------------------------
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from gnn_fpga_b200 import SegmentClassifier, DeviceGraphBatch, _lib
    from gnn_fpga_b200.graph import _ptr, _stream_ptr
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    cfg = WORKLOADS[args.workload]
    graphs = make_graphs(args.workload, rank)
    torch.manual_seed(0)
    masks_e = masks_n = None
    if cfg.get("masked"):
        from gnn_fpga_b200 import data as _data
        masks_e, masks_n = _data.random_masks(cfg["F"], cfg["h"], keep=0.5, seed=1234)
    model = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"], masks_e=masks_e, masks_n=masks_n).to(dev).eval()
    batch = DeviceGraphBatch.from_sparse_graphs(graphs, dev)
    n_real = batch.count_real_edges()
    n_events = len(graphs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # ---- value: batch resident in HBM ----------------------------------------------------
    with torch.no_grad():
        sampler = ClockSampler(local_rank) if rank == 0 else None
        for _ in range(args.warmup):
            model(batch)
        barrier()
        if sampler:
            sampler.mark_start()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.zero_()
            starts[i].record()
            model(batch)
            ends[i].record()
        barrier()
        if sampler:
            sampler.mark_end()
        clocks = sampler.stop() if sampler else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(sum(step_ms))
    # pack x2, input, edge x(it+1), node x it (two kernels per node step at hidden_dim = 32 and 64)
    launches_per_step = 2 + 1 + (cfg["n_iters"] + 1) + cfg["n_iters"] * (2 if cfg["h"] in (32, 64) else 1)

    # ---- per-kernel durations (same process, CUDA events around single launches) ----------
    L = _lib.lib()
    h, F, it = cfg["h"], cfg["F"], cfg["n_iters"]
    blob = model.pack_weights()
    X4 = torch.empty(batch.n_nodes, 4, device=dev)
    Q = [torch.empty(batch.n_nodes, 3 * h, device=dev) for _ in range(2)]
    P = torch.empty(batch.n_nodes, 2 * h, device=dev)
    e = torch.empty(batch.n_slots, device=dev)
    e_in = torch.empty(batch.n_slots, device=dev)
    e_out = torch.empty(batch.n_slots, device=dev)
    split = h in (32, 64)                # node step = gather kernel + tensor-core MLP kernel (include/gnnseg.h)
    h1 = torch.empty(batch.n_nodes, h, device=dev) if split else None
    kt = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rc = fn(); b.record()
        assert rc == 0, (name, rc)
        kt.setdefault(name, []).append((a, b))

    st = _stream_ptr(dev)
    reps = max(3, min(args.steps, 20))
    for rep in range(reps + 1):
        if rep == 1:
            kt.clear()                                      # drop the warm-up pass
        flush.zero_()
        timed("input", lambda: L.gnnseg_input_step(_ptr(blob), _ptr(batch.X), batch.n_nodes, F, h, _ptr(X4), _ptr(P), _ptr(Q[0]), st))
        cur = 0
        for i in range(it):
            qo = _ptr(Q[cur ^ 1]) if i + 1 < it else None
            timed("edge", lambda: L.gnnseg_edge_step(_ptr(blob), C.byref(batch.struct), _ptr(P), h, None, _ptr(e_in), _ptr(e_out), st))
            if split:
                timed("node_gather", lambda: L.gnnseg_node_gather_step(C.byref(batch.struct), _ptr(Q[cur]), _ptr(e_in), _ptr(e_out), h, _ptr(h1), h, st))
                timed("node_mlp", lambda: L.gnnseg_node_mlp_step(_ptr(blob), _ptr(X4), _ptr(h1), h, batch.n_nodes, h, _ptr(P), qo, st))
            else:
                timed("node", lambda: L.gnnseg_node_step(_ptr(blob), C.byref(batch.struct), _ptr(X4), _ptr(Q[cur]), _ptr(e_in), _ptr(e_out), h, _ptr(P), qo, st))
            cur ^= 1
        timed("edge", lambda: L.gnnseg_edge_step(_ptr(blob), C.byref(batch.struct), _ptr(P), h, _ptr(e), None, None, st))
    torch.cuda.synchronize(dev)
    kernel_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in kt.items() if v}
    kernel_share = {k: kernel_ms[k] * {"input": 1, "edge": it + 1}.get(k, it) for k in kernel_ms}

    # ---- e2e: host SparseGraph tuples -> scores on the host ---------------------------------
    e2e = None
    if not args.no_e2e:
        model.use_cuda_graph = False
        host_out = torch.empty((len(graphs), batch.e_max), dtype=torch.float32, pin_memory=True)
        with torch.no_grad():
            for _ in range(3):
                host_out.copy_(model(graphs), non_blocking=True)
                torch.cuda.synchronize(dev)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                host_out.copy_(model(graphs), non_blocking=True)   # scores land in pinned host memory
                torch.cuda.synchronize(dev)                        # every step ends with its result on the host
            dt = time.perf_counter() - t0
        model.use_cuda_graph = True
        h2d = batch.X.numel() * 4 + batch.src.numel() * 4 + batch.dst.numel() * 4
        d2h = host_out.numel() * 4
        e2e = {"sec": dt, "h2d": h2d, "d2h": d2h}
        # the same work, pipelined: predict_stream keeps two batches in flight (host packing of
        # batch i+1 overlaps the GPU work of batch i); every batch still does its own H2D and D2H
        for _ in model.predict_stream([graphs] * 3):
            pass
        barrier()
        t0 = time.perf_counter()
        n_out = 0
        for out_host in model.predict_stream([graphs] * args.steps):
            n_out += 1
        torch.cuda.synchronize(dev)
        e2e["sec_pipelined"] = time.perf_counter() - t0
        assert n_out == args.steps and torch.equal(out_host, host_out)

    # ---- training step (BASELINE configs[4]): forward_train + BCE + L1 + backward + all-reduce + Adam ----
    train = None
    if not args.no_train:
        from gnn_fpga_b200.training import NativeTrainer
        tmodel = SegmentClassifier(cfg["F"], cfg["h"], cfg["n_iters"], masks_e=masks_e, masks_n=masks_n).to(dev).train()
        tmodel.load_state_dict(model.state_dict())
        y = torch.zeros((n_events, batch.e_max), dtype=torch.float32)
        for b_, g_ in enumerate(graphs):
            y[b_, :g_.y.shape[0]] = torch.from_numpy(g_.y)
        y = y.to(dev)
        trainer = NativeTrainer(tmodel, l1=1e-4)
        for _ in range(3):
            trainer.step(batch, y)
        barrier()
        n_tr = max(3, min(args.steps, 20))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_tr):
            trainer.step(batch, y)
        b.record()
        barrier()
        train = {"ms": a.elapsed_time(b) / n_tr, "loss": float(trainer.loss.item())}
        del trainer, tmodel
        batch._ws = {k: v for k, v in batch._ws.items() if not (isinstance(k, tuple) and k[0] == "train")}

    # ---- optional collective: all ranks receive all scores (NCCL all-gather, not in `value`) ------
    gather_ms = None
    if world > 1:
        from gnn_fpga_b200.dist import gather_scores, shard_bounds
        bounds = shard_bounds(world * n_events, world)
        local = batch.scores_2d()
        gather_scores(local, bounds)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            allsc = gather_scores(local, bounds)
        b.record()
        torch.cuda.synchronize(dev)
        gather_ms = a.elapsed_time(b) / 5
        assert allsc.shape[0] == world * n_events

    # ---- reduce over ranks ---------------------------------------------------------------------
    stats = torch.tensor([total_ms, e2e["sec"] if e2e else 0.0, gather_ms or 0.0,
                          e2e["sec_pipelined"] if e2e else 0.0, train["ms"] if train else 0.0], dtype=torch.float64, device=dev)
    counts = torch.tensor([n_real, n_events], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_ms, e2e_sec, gather_ms_max, e2e_pipe_sec, train_ms = stats.tolist()
    all_edges, all_events = counts.tolist()

    if rank == 0:
        ab = algorithmic_bytes(batch.n_nodes, n_real, F, h, it)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        dom = max(kernel_share, key=kernel_share.get)
        dom_bytes = ab[dom]
        try:     # DRAM bytes per launch of that kernel, from the committed ncu capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))[args.workload][dom]
        except Exception:
            traffic = None
        achieved = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9
        ms_per_step = total_ms / args.steps
        fwd_gbs = ab["forward"] / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "segment_classifier_forward_edges_per_sec",
            "value": all_edges * args.steps / (total_ms * 1e-3),
            "unit": "edges/s",
            "events_per_sec": all_events * args.steps / (total_ms * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "hidden_dim": h, "n_iters": it, "batch_per_gpu": n_events,
                       "nodes_per_gpu": batch.n_nodes, "edges_per_gpu": n_real, "parallelism": "events sharded, dp%d" % world,
                       "l2": "flushed between timed steps (256 MiB write)", "cuda_graph": True},
            "roofline": {"bound": "hbm", "kernel": dom + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": kernel_ms[dom]},
            "roofline_forward": {"achieved": fwd_gbs, "frac": fwd_gbs / peak, "unit": "GB/s",
                                 "algorithmic_bytes": ab["forward"]},
            "kernel_ms": kernel_ms, "kernel_ms_per_step": kernel_share,
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = {"value": all_edges * args.steps / e2e_pipe_sec, "unit": "edges/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "ms_per_step": e2e_pipe_sec / args.steps * 1e3,
                           "path": "model.predict_stream(batches of host SparseGraph): per batch C host packing into pinned memory, "
                                   "H2D, device CSR build, forward, D2H into pinned memory; two batches in flight",
                           "unpipelined": {"value": all_edges * args.steps / e2e_sec, "ms_per_step": e2e_sec / args.steps * 1e3,
                                           "path": "model(graphs) then copy to pinned host memory, synchronised every step"}}
        if train:
            n_k = 2 + 1 + (it + 1) + it * (2 if h in (32, 64) else 1) + 2 + 1 + 3 * (it + 1) + 1 + 1 + 1
            line["train_step"] = {"ms": train_ms, "edges_per_sec": all_edges / (train_ms * 1e-3),
                                  "events_per_sec": all_events / (train_ms * 1e-3), "loss": train["loss"],
                                  "gpu_launches_per_step": n_k,
                                  "path": "NativeTrainer.step on the resident batch: gnnseg_forward_train, gnnseg_bce_loss, "
                                          "gnnseg_backward, gnnseg_l1_penalty, %sgnnseg_adam_step; L2 not flushed"
                                          % ("NCCL all-reduce of the flat gradient, " if world > 1 else "")}
        if world > 1:
            line["scores_allgather_ms"] = gather_ms_max    # NCCL all-gather of every rank's (B, E_max) scores
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
