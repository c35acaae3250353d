"""Sharding of independent samples over the GPUs of one box.

One process per GPU (torchrun); samples are independent, so the data path has NO collective: each
rank draws and optimises its contiguous block of the global sample index (the RNG is keyed by the
global index, so results do not depend on the number of ranks).  The only exchanges are the final
gathers of small candidate records.  Works on any torch.distributed backend (nccl on the GPUs,
gloo in the CPU tests); without an initialised process group everything degenerates to one rank.
"""
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank=None, world=None):
    """Contiguous block [first, first+count) of `total` items owned by `rank`; the remainder goes to
    the lowest ranks, so counts differ by at most one."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(int(total), world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def round_robin(n_items, rank=None, world=None):
    """Indices of the items rank `rank` handles when a sorted candidate list is re-dealt
    round-robin (keeps the cheap and the expensive candidates evenly spread)."""
    if rank is None or world is None:
        rank, world = rank_world()
    return list(range(rank, int(n_items), world))


def _comm_device(t):
    """Device collectives must run on: the tensor's own device for nccl, CPU for gloo."""
    if dist.get_backend() == 'nccl':
        return t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def gather_rows(t):
    """Concatenate every rank's [n_r, ...] tensor along dim 0 (n_r may differ per rank, including
    0), in rank order, on every rank.  One all_gather of the sizes plus one of the padded data."""
    rank, world = rank_world()
    if world == 1:
        return t
    dev = _comm_device(t)
    src = t.to(dev).contiguous()
    n = torch.tensor([src.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    if m == 0:
        return t[:0]
    pad = torch.zeros((m,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    pad[:src.shape[0]] = src
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)
    return out.to(t.device)


def gather_round_robin(t, n_items):
    """Inverse of `round_robin`: every rank holds the rows for its dealt indices (in order); returns
    the full [n_items, ...] tensor in the original order on every rank."""
    rank, world = rank_world()
    if world == 1:
        return t
    allrows = gather_rows(t)
    order = [i for r in range(world) for i in range(r, int(n_items), world)]
    out = torch.empty_like(allrows)
    out[torch.tensor(order, dtype=torch.int64, device=allrows.device)] = allrows
    return out


def all_sum(x):
    """Sum of a python number over ranks."""
    rank, world = rank_world()
    if world == 1:
        return x
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    return t.item()
