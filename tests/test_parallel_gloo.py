"""N > 1 host logic of the data-parallel path on CPU: world_size-2 gloo.

Covers motion_style_transfer_b200/parallel.py: contiguous agent sharding (ragged and empty shards), the
end-of-scene row gather, the (sum ADE, sum FDE, n) reduce, and the single flat-gradient all-reduce whose rank-mean
must equal the full-batch gradient when per-rank losses are weighted by shard size (utils/train_epoch.py).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motion_style_transfer_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, out_q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        out_q.put((rank, fn(rank, world)))
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return res


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 2, 7, 10, 64, 1001):
        for w in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(w):
                lo, hi = parallel.shard_bounds(n, r, w)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
            sizes = parallel.shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)
    assert parallel.world() == (0, 1)
    t = torch.arange(6.).reshape(3, 2)
    assert parallel.gather_rows(t, 3) is t                      # single process: no copy, no collective
    assert parallel.reduce_metric_sums(6.0, 3.0, 3) == (2.0, 1.0, 3)


def _gather_case(rank, world):
    out = {}
    for n in (7, 1, 0, 4):                                         # ragged, fewer agents than ranks, empty, even
        lo, hi = parallel.shard_bounds(n)
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 1.5
        got = parallel.gather_rows(full[lo:hi].contiguous(), n)
        out[n] = torch.equal(got, full)
    lo, hi = parallel.shard_bounds(7)
    ade = torch.arange(7, dtype=torch.float32)
    out['metrics'] = parallel.reduce_metric_sums(ade[lo:hi].sum().item(), 2 * ade[lo:hi].sum().item(), hi - lo)
    return out


def test_gather_rows_and_metric_reduce_gloo():
    res = _run(_gather_case)
    for r in (0, 1):
        assert all(res[r][n] for n in (7, 1, 0, 4))
        assert res[r]['metrics'] == (3.0, 6.0, 7)


def _grad_case(rank, world):
    """Rank-mean of shard-weighted local gradients == gradient of the full-batch mean loss (ragged shards)."""
    torch.manual_seed(0)
    Bn = 5
    x = torch.randn(Bn, 4)
    y = torch.randn(Bn, 3)
    w1 = torch.nn.Parameter(torch.randn(4, 3) * 0.3)
    w2 = torch.nn.Parameter(torch.randn(3) * 0.1)
    frozen = torch.nn.Parameter(torch.randn(2), requires_grad=False)
    full = ((x @ w1 + w2 - y) ** 2).mean()
    g1, g2 = torch.autograd.grad(full, [w1, w2])
    lo, hi = parallel.shard_bounds(Bn)
    local = ((x[lo:hi] @ w1 + w2 - y[lo:hi]) ** 2).mean() * ((hi - lo) * world / Bn)
    local.backward()
    flat, ps = parallel.flatten_grads([w1, frozen, w2])
    scale = parallel.allreduce_flat(flat)
    flat = flat * scale
    ok = len(ps) == 2 and flat.numel() == 15
    ok = ok and torch.allclose(flat[:12].reshape(4, 3), g1, atol=1e-6) and torch.allclose(flat[12:], g2, atol=1e-6)
    # broadcast_params: every rank ends up with rank 0's adapter init
    p = torch.nn.Parameter(torch.full((3,), float(rank + 1)))
    parallel.broadcast_params([p])
    return bool(ok and torch.equal(p.data, torch.ones(3)))


def test_flat_gradient_allreduce_matches_full_batch_gloo():
    res = _run(_grad_case)
    assert res[0] and res[1]


def _layout_case(rank, world):
    """Fixed-layout flat gradients: a tensor without a gradient on one rank still lines up, the loss mean
    over ranks is sum / world (count 1 per rank), and the shared generator gives every rank the same scene order."""
    a = torch.nn.Parameter(torch.ones(3))
    b = torch.nn.Parameter(torch.ones(2))
    if rank == 0:
        a.grad = torch.full((3,), 2.0)          # rank 0: only `a` has a gradient
    else:
        a.grad = torch.full((3,), 4.0)
        b.grad = torch.full((2,), 6.0)
    flat, ps = parallel.flatten_grads([a, b], fixed_layout=True)
    ok = len(ps) == 2 and flat.numel() == 5
    scale = parallel.allreduce_flat(flat)
    ok = ok and torch.allclose(flat * scale, torch.tensor([3.0, 3.0, 3.0, 3.0, 3.0]))
    mean_loss, _, n = parallel.reduce_metric_sums(10.0 * (rank + 1), 0.0, 1)
    ok = ok and abs(mean_loss - 15.0) < 1e-12 and n == world
    torch.manual_seed(100 + rank)               # different default streams per process
    perm = torch.randperm(16, generator=parallel.shared_generator()).tolist()
    parallel.barrier()
    return bool(ok), perm


def test_fixed_layout_loss_mean_and_shared_generator_gloo():
    res = _run(_layout_case)
    assert res[0][0] and res[1][0]
    assert res[0][1] == res[1][1]


def _init_env_worker(rank, world, port, out_q):
    """What torchrun gives a rank: only environment variables.  (Not via _worker: that one creates the group itself.)"""
    import io
    import sys
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    try:
        from motion_style_transfer_b200 import parallel
        sys.stdout = buf = io.StringIO()
        got = parallel.init_from_env()
        muted = sys.stdout is not buf
        again = parallel.init_from_env()                      # a second call joins nothing and changes nothing
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t)
        out_q.put((rank, dict(world=tuple(got), again=tuple(again), muted=muted, total=float(t), backend=dist.get_backend())))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        out_q.put((rank, repr(e)))
        raise


def test_init_from_env_joins_the_torchrun_group_gloo():
    """The entry points (train / test / evaluate_multickpts) call parallel.init_from_env(): under torchrun it forms the group
    from the environment (gloo here, NCCL with CUDA) and silences the ranks other than 0; a plain launch is untouched."""
    from motion_style_transfer_b200 import parallel
    assert parallel.init_from_env() == (0, 1) and not dist.is_initialized()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_init_env_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert not isinstance(res[r], str), res[r]
        assert res[r]['world'] == (r, 2) and res[r]['again'] == (r, 2) and res[r]['total'] == 3.0 and res[r]['backend'] == 'gloo'
    assert res[0]['muted'] is False and res[1]['muted'] is True
