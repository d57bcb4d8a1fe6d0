"""Event sharding + score gathering on CPU (gloo, world_size 2): the host logic of the
multi-GPU path.  The per-rank forward is replaced by the CPU oracle here (tests may use it);
on GPUs the same code runs with the CUDA model and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_fpga_b200 import data
from gnn_fpga_b200.dist import shard_bounds, sharded_predict
from oracle import segclf_oracle as O


def test_shard_bounds_cover_and_balance():
    for n, w in ((64, 8), (10, 4), (3, 8), (0, 2), (7, 1)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    wts = [100, 1, 1, 1, 100, 1, 1, 100]
    b = shard_bounds(8, 3, wts)
    assert b[0][0] == 0 and b[-1][1] == 8 and all(b[i][1] == b[i + 1][0] for i in range(2))
    loads = [sum(wts[lo:hi]) for lo, hi in b]
    assert max(loads) <= 1.5 * sum(wts) / 3
    with pytest.raises(ValueError):
        shard_bounds(4, 2, [1, 2, 3])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_forward(params, n_iters):
    def fwd(graphs):
        X, src, dst, e_max = O.flatten_sparse_batch(graphs)
        return O.sparse_forward(params, X, src, dst, n_iters).reshape(len(graphs), e_max)
    return fwd


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    graphs = [data.acts_like_graph(n, seed=i) for i, n in enumerate((12, 20, 9, 15, 11))]
    params = O.init_params(3, 8, seed=0)
    scores, bounds = sharded_predict(_oracle_forward(params, 2), graphs)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), scores.numpy())
    np.save(os.path.join(out_dir, "b%d.npy" % rank), np.array(bounds))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_predict_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1, equal_nan=True)          # every rank holds all scores
    graphs = [data.acts_like_graph(n, seed=i) for i, n in enumerate((12, 20, 9, 15, 11))]
    params = O.init_params(3, 8, seed=0)
    fwd = _oracle_forward(params, 2)
    for b, g in enumerate(graphs):
        n_e = g.Ri_rows.shape[0]
        one = fwd([g])[0].numpy()
        # same as the single-event result: an event's scores do not depend on its batch.  (The
        # torch-CPU oracle's BLAS blocks differently per batch size, hence a 1e-6 tolerance here;
        # the CUDA kernels are bit-equal, see test_gpu_parity.py::test_events_are_independent.)
        assert np.allclose(r0[b, :n_e], one[:n_e], rtol=1e-6, atol=0)
        assert np.all(np.isnan(r0[b, n_e:]) | (r0[b, n_e:] > 0))
    bounds = np.load(tmp_path / "b0.npy")
    assert bounds[0][0] == 0 and bounds[-1][1] == len(graphs)


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnn_fpga_b200.training import allreduce_gradients
    torch.manual_seed(0)
    lin = [torch.nn.Linear(3, 4), torch.nn.Linear(4, 1)]
    params = [p for l in lin for p in l.parameters()]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    params[1].grad = None                       # a parameter without gradient is skipped
    allreduce_gradients(params)
    np.save(os.path.join(out_dir, "g%d.npy" % rank),
            np.concatenate([p.grad.reshape(-1).numpy() for p in params if p.grad is not None]))
    dist.barrier()
    dist.destroy_process_group()


def _weighted_grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnn_fpga_b200.training import allreduce_gradients
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.full((3,), float(rank + 1))
    allreduce_gradients([p], n_local=(30 if rank == 0 else 10))       # uneven shards: 30 and 10 padded slots
    np.save(os.path.join(out_dir, "w%d.npy" % rank), p.grad.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_weights_uneven_shards(tmp_path):
    """Uneven shards: the ranks' mean-over-local-slots gradients are weighted by their slot counts, which is
    the gradient of the reference's mean over ALL padded slots (gnn/estimator.py:57) of the union batch."""
    mp.spawn(_weighted_grad_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "w0.npy"), np.load(tmp_path / "w1.npy")
    assert np.array_equal(g0, g1) and np.allclose(g0, (30 * 1.0 + 10 * 2.0) / 40)


def _gatherer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gnn_fpga_b200.dist import ScoreGatherer
    g = ScoreGatherer()
    outs = []
    for step, (e0, e1) in enumerate(((5, 7), (5, 7), (4, 6))):                  # same bucket twice, then a narrower one
        E = e0 if rank == 0 else e1
        local = torch.full((2, E), float(10 * step + rank))
        if step == 2:
            g.reset()
        outs.append(g(local).clone().numpy())
    np.save(os.path.join(out_dir, "s%d.npy" % rank), np.stack([np.pad(o, ((0, 0), (0, 7 - o.shape[1])), constant_values=-1) for o in outs]))
    dist.barrier()
    dist.destroy_process_group()


def test_score_gatherer_reuses_buffers_and_keeps_padding_nan(tmp_path):
    """ScoreGatherer: one all_gather_into_tensor per call into a reused buffer; equal B on every rank gives the
    (world * B, E_max) tensor directly, NaN where a rank's block is narrower, also after the bucket changed."""
    mp.spawn(_gatherer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "s0.npy"), np.load(tmp_path / "s1.npy")
    assert np.array_equal(s0, s1, equal_nan=True)
    for step, (e0, e1) in enumerate(((5, 7), (5, 7), (4, 6))):
        o = s0[step][:, :max(e0, e1)]
        assert np.all(o[:2, :e0] == 10 * step) and np.all(np.isnan(o[:2, e0:]))
        assert np.all(o[2:, :e1] == 10 * step + 1)


def test_two_rank_gradient_allreduce_averages(tmp_path):
    """allreduce_gradients: one flat all-reduce, mean over ranks, identical on every rank
    (the NCCL step of the sharded training_step; gloo here)."""
    mp.spawn(_grad_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1)
    # ranks held (i+1) and 2(i+1): the mean is 1.5 (i+1) for parameter i in {0, 2, 3}
    expect = np.concatenate([np.full(12, 1.5), np.full(4, 4.5), np.full(1, 6.0)])
    assert np.allclose(g0, expect)
