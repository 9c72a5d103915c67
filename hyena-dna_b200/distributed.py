"""Data-parallel plumbing for the operator: one process per GPU, batch-sharded replicas.

The (batch x channel) rows of the long convolution are independent, so the path shards over the
batch exactly as the reference's only multi-GPU mode does (Lightning DDP, train.py:612-621): no
activation ever crosses a GPU; the only collective is one all-reduce (sum) of the operator's
parameter gradients per step (~1.2 MB at D=256).  Works with the ``nccl`` backend on GPUs and the
``gloo`` backend on CPU (used by the host-logic tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, world_size, rank):
    """Contiguous batch slice [lo, hi) owned by ``rank``; earlier ranks take the remainder."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, world_size=None, rank=None):
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return x[lo:hi]


def allreduce_grads(params, group=None, average=False):
    """Sum (or average) the gradients of ``params`` across ranks with ONE flat all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()


def allreduce_tensors(tensors, group=None):
    """In-place sum of a list of tensors across ranks with one flat all-reduce."""
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return flat.numel()


def gather_outputs(y, group=None):
    """All-gather batch shards of an output (verification / serving only: 1 GB per sample at L=1M)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return y
    outs = [torch.empty_like(y) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, y.contiguous(), group=group)
    return torch.cat(outs, dim=0)
