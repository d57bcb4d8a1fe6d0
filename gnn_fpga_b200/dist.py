"""Event sharding over the GPUs of one box (one process per GPU, torch.distributed).

Events are independent (the batch is block diagonal), so the forward needs
no exchange step: every rank runs the whole model on its contiguous slice of the event list.
The only collective is an optional all-gather of the (B_local, E_max) score blocks, NCCL on
GPUs (gloo in the CPU tests of the host logic).  Scores are bit-identical to a 1-GPU run
because an event's scores do not depend on the rest of the batch.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_events, world, weights=None):
    """Contiguous split of range(n_events) into `world` slices -> list of (lo, hi).
    With `weights` (e.g. edges per event) the cut points balance the prefix sums, otherwise
    sizes differ by at most one."""
    if world < 1 or n_events < 0:
        raise ValueError("bad shard request")
    if weights is None:
        base, extra = divmod(n_events, world)
        sizes = [base + (1 if r < extra else 0) for r in range(world)]
        cuts = np.concatenate([[0], np.cumsum(sizes)])
    else:
        w = np.asarray(weights, dtype=np.float64)
        if w.shape[0] != n_events:
            raise ValueError("weights must have one entry per event")
        prefix = np.concatenate([[0.0], np.cumsum(w)])
        targets = prefix[-1] * np.arange(1, world) / world
        inner = np.searchsorted(prefix, targets, side="left")
        # choose the closer of the two neighbouring cut points, keep cuts monotone
        inner = np.array([i if i == 0 or abs(prefix[i] - t) <= abs(prefix[i - 1] - t) else i - 1
                          for i, t in zip(inner, targets)], dtype=np.int64)
        cuts = np.concatenate([[0], np.maximum.accumulate(np.clip(inner, 0, n_events)), [n_events]])
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def bind_to_local_cpus(device_index):
    """Restrict this process to the CPU cores next to its GPU (the `local_cpulist` of the GPU's PCI device in sysfs)
    BEFORE it allocates pinned host memory: the event store's arena and the pinned result buffers are then first
    touched, i.e. placed, on the GPU's own NUMA node, and the H2D / D2H copies of several ranks sharing a host do
    not cross the socket interconnect.  Returns the core list, or None when the topology cannot be read (then
    nothing is changed).  One process per GPU (torchrun): call it right after choosing the device."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


class ScoreGatherer:
    """All-gather of the per-rank (B_r, E_r) score blocks into one (sum B_r, E_max) tensor on every rank
    (NaN where an event has fewer than E_max slots on its rank), the only collective of the inference path.

    One `all_gather_into_tensor` per call into a preallocated buffer, enqueued on the current (compute) stream
    with no host synchronisation: the block shapes are exchanged ONCE per shape bucket (`shapes` may also be
    passed in when the caller knows them, e.g. from the event list), not per call.  Ranks must call it in
    step; a rank whose local shape changes triggers a new exchange on all ranks only if every rank passes
    through `exchange_shapes` again, so callers with ragged batches pass `shapes` explicitly or call
    `reset()` on every rank between buckets."""

    def __init__(self, group=None):
        self.group = group
        self.shapes = None
        self._block = None
        self._out = None
        self._dirty_cols = 0

    def reset(self):
        self.shapes = None

    def exchange_shapes(self, local_scores):
        world = dist.get_world_size(self.group)
        shape = torch.tensor([local_scores.shape[0], local_scores.shape[1]], dtype=torch.int64, device=local_scores.device)
        allsh = torch.empty(2 * world, dtype=torch.int64, device=local_scores.device)
        dist.all_gather_into_tensor(allsh, shape, group=self.group)
        v = allsh.tolist()                                   # the one host synchronisation, once per bucket
        self.shapes = [(int(v[2 * r]), int(v[2 * r + 1])) for r in range(world)]
        return self.shapes

    def __call__(self, local_scores, shapes=None):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        if shapes is not None:
            self.shapes = [tuple(int(x) for x in sh) for sh in shapes]
        if self.shapes is None or self.shapes[rank] != tuple(local_scores.shape):
            self.exchange_shapes(local_scores)
        b_max = max(sh[0] for sh in self.shapes)
        e_max = max(sh[1] for sh in self.shapes)
        dev, dt = local_scores.device, local_scores.dtype
        if self._out is None or self._out.shape != (world * b_max, e_max) or self._out.device != dev or self._out.dtype != dt:
            self._out = torch.empty((world * b_max, e_max), dtype=dt, device=dev)
            self._block = torch.full((b_max, e_max), float("nan"), dtype=dt, device=dev)
            self._dirty_cols = 0
        if tuple(local_scores.shape) == (b_max, e_max) and local_scores.is_contiguous():
            block = local_scores                            # the common case needs no staging copy
        else:
            B, E = local_scores.shape
            if E < self._dirty_cols or B < b_max:
                self._block.fill_(float("nan"))             # a narrower block than last time: clear the stale part
            self._block[:B, :E] = local_scores
            self._dirty_cols = E
            block = self._block
        dist.all_gather_into_tensor(self._out, block, group=self.group)
        if all(sh[0] == b_max for sh in self.shapes):
            return self._out                                 # (world * B, E_max): rows already in event order
        return torch.cat([self._out[r * b_max:r * b_max + sh[0]] for r, sh in enumerate(self.shapes)], dim=0)


_default_gatherers = {}


def gather_scores(local_scores, bounds, group=None, shapes=None):
    """All-gather per-rank (B_r, E_r) score blocks into one (B, E_max) tensor on every rank, padded with
    NaN where an event has fewer than E_max slots on its rank (see ScoreGatherer; one gatherer is kept per
    group, so repeated calls reuse its buffers and exchange shapes only when the local shape changes).
    The returned tensor is overwritten by the next call."""
    g = _default_gatherers.setdefault(id(group) if group is not None else None, ScoreGatherer(group))
    out = g(local_scores, shapes)
    assert out.shape[0] == bounds[-1][1]
    return out


def sharded_predict(forward_fn, graphs, weights="edges", gather=True, group=None):
    """Run `forward_fn(list_of_graphs) -> (B_local, E_local)` on this rank's slice of `graphs`
    and (optionally) gather all scores on every rank.  `forward_fn` is the model."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    w = [g.Ri_rows.shape[0] for g in graphs] if weights == "edges" else weights
    bounds = shard_bounds(len(graphs), world, w)
    lo, hi = bounds[rank]
    if hi > lo:
        local = forward_fn(graphs[lo:hi])
    else:
        local = None
    if world == 1 or not gather:
        return local, bounds
    if local is None:
        ref_dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        local = torch.empty((0, 0), dtype=torch.float32, device=ref_dev)
    return gather_scores(local, bounds, group), bounds
