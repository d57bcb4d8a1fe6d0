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


def gather_scores(local_scores, bounds, group=None):
    """All-gather per-rank (B_r, E_r) score blocks into one (B, E_max) tensor on every rank,
    padded with NaN where an event has fewer than E_max slots on its rank.  Blocks are padded
    to a common shape first so a single fixed-size all_gather does the job."""
    world = dist.get_world_size(group)
    dev = local_scores.device
    shape = torch.tensor([local_scores.shape[0], local_scores.shape[1]], dtype=torch.int64, device=dev)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    shapes = [tuple(int(v) for v in s.tolist()) for s in shapes]
    b_max = max(s[0] for s in shapes)
    e_max = max(s[1] for s in shapes)
    block = torch.full((b_max, e_max), float("nan"), dtype=local_scores.dtype, device=dev)
    block[:local_scores.shape[0], :local_scores.shape[1]] = local_scores
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    out = torch.cat([blk[:s[0]] for blk, s in zip(blocks, shapes)], dim=0)
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
