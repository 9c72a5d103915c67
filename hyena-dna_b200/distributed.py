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


class OverlappedGradReducer:
    """Sum the operator's parameter gradients across ranks, overlapped with the tail of the backward pass.

    The gradients of one operator become available in two groups: projections, short filter and filter bias right after
    the fused core's backward node; the implicit-filter MLP (seven small tensors + freq) only after the filter backward
    that follows it (~4 ms later at L = 2^20).  Each group lives in ONE persistent flat buffer (no torch.cat, no
    per-tensor copy kernels: `torch._foreach_copy_`), and the first group's all-reduce is issued on a side stream by a
    post-accumulate-grad hook as soon as its last gradient lands, so it runs under the filter backward.  `finish()`
    reduces the second (35 KB) group, waits for both and hands the sums back to `p.grad`.

    Usage:   red = OverlappedGradReducer(params);  loss.backward();  red.finish()
    Works with NCCL on GPUs (side stream) and with gloo on CPU (synchronous; used by the host-logic tests)."""

    def __init__(self, params, late=lambda name: "implicit_filter" in name, named=None, group=None):
        self.group = group
        named = list(named) if named is not None else [(str(i), p) for i, p in enumerate(params)]
        self.buckets = [[p for n, p in named if p.requires_grad and not late(n)],
                        [p for n, p in named if p.requires_grad and late(n)]]
        self.flat, self.views, self.ready, self.work = [], [], [0, 0], [None, None]
        for bucket in self.buckets:
            n = sum(p.numel() for p in bucket)
            dev = bucket[0].device if bucket else torch.device("cpu")
            flat = torch.zeros(n, dtype=torch.float32, device=dev)
            views, off = [], 0
            for p in bucket:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self.flat.append(flat)
            self.views.append(views)
        self.cuda = bool(self.flat[0].is_cuda) if self.flat[0].numel() else False
        self.side = torch.cuda.Stream(device=self.flat[0].device) if self.cuda else None
        self._handles = []
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _make_hook(self, bi):
        def hook(param):
            self.ready[bi] += 1
            if self.ready[bi] == len(self.buckets[bi]):
                self._launch(bi)
        return hook

    def _launch(self, bi):
        if not self.buckets[bi] or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        torch._foreach_copy_(self.views[bi], [p.grad for p in self.buckets[bi]])
        if self.cuda:
            self.side.wait_stream(torch.cuda.current_stream(self.flat[bi].device))
            with torch.cuda.stream(self.side):
                self.work[bi] = dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """Call after backward(): gradients of every parameter hold the sum over ranks on return (stream-ordered)."""
        active = dist.is_initialized() and dist.get_world_size(self.group) > 1
        for bi, bucket in enumerate(self.buckets):
            if bucket and self.ready[bi] != len(bucket) and self.ready[bi] != 0:
                raise RuntimeError("OverlappedGradReducer: some parameters of a group received no gradient")
            if active and bucket and self.ready[bi] == len(bucket):
                if self.work[bi] is not None:
                    self.work[bi].wait()                      # makes the current stream wait for the collective
                if self.cuda:
                    torch.cuda.current_stream(self.flat[bi].device).wait_stream(self.side)
                torch._foreach_copy_([p.grad for p in bucket], self.views[bi])
            self.ready[bi] = 0
            self.work[bi] = None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


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
