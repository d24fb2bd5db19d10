"""Data-parallel plumbing of the forecasting path: one process per GPU, agents sharded, no data-path collective.

The reference is single-process (trainer.py:54-57); agents of a scene are independent (evaluate.py:109), so
ranks take contiguous agent ranges of every scene and only two tiny exchanges exist:

  * evaluation -- per-agent (ade, fde) rows are gathered once at the end so that every rank returns the same
    ``(ade, fde, df_out)`` the single-process call would (evaluate.py:297-307);
  * fine-tuning -- the flattened gradient of the trainable (LoRA) tensors, 8 190 floats for Y-Net mosa_1, is summed
    with ONE all-reduce per optimiser step and averaged inside the Adam kernel (``models.trainer.FusedAdam``).

Everything here is device-agnostic host logic (tensors stay where they are; NCCL on the GPU box, gloo in the CPU
tests).  Nothing in this module touches the CUDA library.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(quiet=True):
    """Join the process group ``torchrun`` describes (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* in the environment): NCCL with
    one GPU per process, gloo without CUDA.  A no-op for a plain ``python`` launch or when a group exists already.  With
    ``quiet`` the ranks other than 0 stop printing: every rank computes the same metrics, and the log scraper
    (utils/extract_log.py) expects one run per parameter dictionary in a log.  Returns (rank, world_size)."""
    import os
    import sys
    n = int(os.environ.get('WORLD_SIZE', '1'))
    if n > 1 and dist.is_available() and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if torch.cuda.is_available():
            local = int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0')))
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group('gloo')
        if quiet and dist.get_rank() != 0:
            sys.stdout = open(os.devnull, 'w')
    return world()


def shard_bounds(n, rank=None, world_size=None):
    """Contiguous [lo, hi) of ``n`` agents owned by ``rank``: the first ``n % world`` ranks get one extra agent.

    Contiguity keeps the gathered rows in the reference's agent order (evaluate.py:109 walks agents in order).
    """
    if rank is None or world_size is None:
        rank, world_size = world()
    if n < 0 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f'bad shard request n={n} rank={rank} world={world_size}')
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world_size):
    return [shard_bounds(n, r, world_size)[1] - shard_bounds(n, r, world_size)[0] for r in range(world_size)]


def gather_rows(local, n_total):
    """Concatenate per-rank row blocks (contiguous shards of ``n_total`` rows, see ``shard_bounds``) on every rank.

    ``local``: (n_local, ...) tensor.  Ragged shards are padded to the largest one for the all-gather.
    """
    rank, ws = world()
    if ws == 1:
        return local
    sizes = shard_sizes(n_total, ws)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f'rank {rank} holds {local.shape[0]} rows, its shard of {n_total} is {sizes[rank]}')
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def reduce_metric_sums(sum_ade, sum_fde, count, device=None):
    """One SUM all-reduce of (sum ADE, sum FDE, n) -> global means.  ``count`` is THIS rank's share of n
    (agents of its shard; 1 per rank for a mean over ranks): the counts are summed like the values."""
    t = torch.tensor([float(sum_ade), float(sum_fde), float(count)], dtype=torch.float64, device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    n = max(t[2].item(), 1.0)
    return t[0].item() / n, t[1].item() / n, int(t[2].item())


def flatten_grads(params, fixed_layout=False):
    """The trainable tensors' gradients as ONE contiguous buffer (+ the params that contributed, in order).

    ``fixed_layout`` (used whenever the buffer feeds a collective): every trainable tensor takes part, a missing
    gradient counts as zero -- all ranks then build the same layout even if one of them saw no agents or a tensor
    received no gradient on it, so the all-reduce sizes always match."""
    if fixed_layout:
        ps = [p for p in params if p.requires_grad]
        if not ps:
            return None, ps
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
        return flat, ps
    ps = [p for p in params if p.requires_grad and p.grad is not None]
    if not ps:
        return None, ps
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    return flat, ps


def allreduce_flat(flat):
    """SUM over ranks, in place; returns the factor the consumer must apply to get the mean (1 / world)."""
    ws = world()[1]
    if ws > 1 and flat is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / ws


def broadcast_params(params, src=0):
    """Make every rank start from rank ``src``'s trainable tensors (LoRA A is random-initialised per process)."""
    if world()[1] == 1:
        return
    for p in params:
        dist.broadcast(p.data, src=src)


def barrier():
    if world()[1] > 1:
        dist.barrier()


def shared_generator(seed=None):
    """A torch.Generator whose state is identical on every rank (rank 0 draws the seed, everyone adopts it).

    ``train_epoch`` shards the AGENTS of each batch across ranks, which presumes that all ranks walk the scenes in the
    same order: the shuffling train DataLoader must therefore draw from a shared stream, not from each process's own
    default generator."""
    if seed is None:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    if world()[1] > 1:
        t = torch.tensor([seed], dtype=torch.int64)
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.broadcast(t, src=0)
        seed = int(t.item())
    g = torch.Generator()
    g.manual_seed(seed)
    return g
