"""batch_generator: the reference's generator protocol (gnn/trainSegmentClassifier.py:97-111) on the
GPU path.

The reference's generator yields, for ever, `(batch_inputs, batch_target)` with
`batch_inputs = [X, Ri, Ro]` dense CUDA tensors it builds on the CPU for every batch
(graph_from_sparse + merge_graphs + np_to_torch), and `Estimator.fit_gen` / `Estimator.predict`
(gnn/estimator.py:80-146) call `next(generator)` once per batch.  This generator keeps the protocol
-- same arguments, same batch cuts (`graphs[j : j + batch_size]`), same padded `(B, E_max)` float32
target with zeros in the padding -- but `batch_inputs` is a `DeviceGraphBatch` (the sparse device form
`SegmentClassifier.forward` takes; the dense tensors would be 41 GB for 64 ACTS events) assembled from a
`GraphStore`: no per-batch host work, and the copies of the next batches are already in flight on a side
stream while the caller works on the current one.
"""
import torch

from .graph import DeviceBatchBuffers, DeviceGraphBatch, _require_cuda
from .store import GraphStore


def batch_generator(graphs, n_samples=1, batch_size=1, train=True, device="cuda", depth=2, reorder="auto"):
    """`graphs`: a list of SparseGraph tuples (as load_graphs(filenames, SparseGraph) returns) or a GraphStore
    built from it once.  Yields `(DeviceGraphBatch, target)` for ever, epoch after epoch, like the reference.
    `train` is accepted for signature compatibility (it only set `volatile` in the reference, a no-op on
    torch >= 0.4).  A yielded batch stays valid until `depth` further batches have been requested."""
    dev = _require_cuda(torch.device(device))
    store = graphs if isinstance(graphs, GraphStore) else GraphStore.from_sparse_graphs(list(graphs)[:n_samples], reorder=reorder)
    batches = store.batches(batch_size, n_samples)
    if not batches:
        raise ValueError("no graphs to batch")
    depth = max(1, int(depth))
    n, n_in, n_out, n_slots, B = store.max_batch_shape(batches)
    slots = [{"bufs": DeviceBatchBuffers(dev, n, n_in, n_out, n_slots, B, store.F, store.col_bytes),
              "target": torch.empty(max(n_slots, 1), dtype=torch.float32, device=dev), "free": None}
             for _ in range(depth + 1)]
    targets_host = {}                                    # padded labels per batch, built once (pinned)
    side = torch.cuda.Stream(dev)

    def issue(k):
        """Copies + device assembly of batch k on the side stream, into slot k % (depth + 1)."""
        sb = batches[k % len(batches)]
        s = slots[k % len(slots)]
        j = k % len(batches)
        if j not in targets_host:
            targets_host[j] = store.targets(sb.lo, sb.hi)
        th = targets_host[j]
        with torch.cuda.stream(side):
            if s["free"] is not None:
                side.wait_event(s["free"])               # the caller is done with the slot's previous batch
            batch = DeviceGraphBatch.from_store(store, sb.lo, sb.hi, dev, bufs=s["bufs"])
            target = s["target"][:th.numel()].view(th.shape)
            target.copy_(th, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(side)
        return batch, target, ready, s

    k = 0
    ahead = [issue(i) for i in range(depth)]
    while True:
        batch, target, ready, s = ahead.pop(0)
        torch.cuda.current_stream(dev).wait_event(ready)
        yield batch, target
        # the caller asked for the next batch: everything it queued on this one is on its stream by now
        s["free"] = torch.cuda.Event()
        s["free"].record(torch.cuda.current_stream(dev))
        ahead.append(issue(k + depth))
        k += 1
