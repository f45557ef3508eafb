"""Image-parallel sharding of the hot path: one process per GPU, no data-path collective.

Every rank extracts the (replicated) prior shape and renders its own contiguous slice of the global batch - exactly
what the reference gets from accelerate/DDP (Trainer.py:170-180).  The only exchange is the all-reduce of parameter
gradients after the backward; nothing in libb2a.so communicates.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous [start, stop) slice of a global batch for `rank` (accelerate's default split: equal shards, the first
    `n_items % world_size` ranks take one extra item)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_gradients(tensors, average=True, group=None):
    """DDP semantics on a list of gradient tensors (None entries are skipped): one flat bucket, one all-reduce, copied
    back in place.  Returns the number of bytes reduced.  No-op outside a process group."""
    grads = [t for t in tensors if t is not None]
    if not grads or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    if len(grads) == 1 and grads[0].is_contiguous():
        # single bucket already: reduce in place (no flatten / copy-back launches)
        if average and dist.get_backend(group) == "nccl":
            dist.all_reduce(grads[0], op=dist.ReduceOp.AVG, group=group)    # NCCL averages inside the collective: no div launch
        else:
            dist.all_reduce(grads[0], op=dist.ReduceOp.SUM, group=group)
            if average:
                grads[0].div_(world)
        return grads[0].numel() * grads[0].element_size()
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()
