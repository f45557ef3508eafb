"""Image-parallel sharding of the hot path: one process per GPU, no data-path collective.

Every rank extracts the (replicated) prior shape and renders its own contiguous slice of the global batch - exactly
what the reference gets from accelerate/DDP (Trainer.py:170-180).  The only exchange is the all-reduce of parameter
gradients after the backward; nothing in libb2a.so communicates.
"""
import os

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


class GradientBuckets:
    """DDP-style bucketed gradient all-reduce, overlapped with the backward pass (reference: accelerate / DistributedDataParallel,
    Trainer.py:170-180; SURVEY.md §2.2 sizes the MagicPony message at ~68 MB of fp32 parameter gradients per step).

    Flat fp32 buckets of at most `bucket_bytes`; `launch(i)` enqueues bucket i's all-reduce (NCCL AVG) on a SIDE stream behind an
    event recorded on the compute stream, so the collective runs under whatever backward work is still queued; `wait()` makes the
    compute stream wait for every launched bucket - call it where the optimiser would read the gradients.  Outside a process
    group (N = 1) every method is a no-op.  Nothing here touches libb2a.so: the path itself has no exchange step (§8e)."""

    def __init__(self, total_bytes, device, bucket_bytes=25 << 20, group=None, tail_bytes=0):
        """tail_bytes > 0: the LAST bucket holds exactly that many bytes (the gradients that become final last, e.g. d_sdf) and
        the rest of the set is cut into `bucket_bytes` buckets in front of it - the exposed collective is as small as possible."""
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        n = max(int(total_bytes) // 4, 1)
        per = max(int(bucket_bytes) // 4, 1)
        tail = min(max(int(tail_bytes) // 4, 0), n)
        self.flat = torch.zeros(n, device=device)
        self.buckets = [self.flat[i:min(i + per, n - tail)] for i in range(0, n - tail, per)] + ([self.flat[n - tail:]] if tail else [])
        self.stream = torch.cuda.Stream(device=device) if (self.active and torch.device(device).type == "cuda") else None
        self.ready = [torch.cuda.Event() for _ in self.buckets] if self.stream is not None else []
        self.launched = []
        self.bytes_reduced = 0

    def view(self, bucket, numel, offset=0):
        """A [numel] slice of a bucket: point a gradient at it so that the kernel writes straight into the bucket (no copy)."""
        return self.buckets[bucket][offset:offset + numel]

    def launch(self, i):
        if not self.active:
            return
        b = self.buckets[i]
        op = dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM
        if self.stream is None:       # gloo / CPU: synchronous
            dist.all_reduce(b, op=op, group=self.group)
            if op == dist.ReduceOp.SUM:
                b.div_(dist.get_world_size(self.group))
        else:
            self.ready[i].record()                    # the bucket's gradients are final at this point of the compute stream
            self.stream.wait_event(self.ready[i])
            with torch.cuda.stream(self.stream):
                dist.all_reduce(b, op=op, group=self.group)
        self.launched.append(i)
        self.bytes_reduced += b.numel() * 4

    def launch_many(self, indices):
        """Buckets `indices` in ONE enqueue: the collectives are issued inside one NCCL group (c10d coalescing manager), i.e. one
        launch on the side stream instead of one c10d call + launch per bucket - what matters on a host-bound rank.  (Capturing the
        collectives into a CUDA graph was tried and deadlocked on the 2-GPU box; not used.)  Falls back to launch() per bucket."""
        indices = tuple(indices)
        if not self.active or not indices:
            return
        if self.stream is None or len(indices) == 1 or os.environ.get("B2A_ALLREDUCE_COALESCE", "1") == "0":
            for i in indices:
                self.launch(i)
            return
        self.ready[indices[0]].record()
        self.stream.wait_event(self.ready[indices[0]])
        with torch.cuda.stream(self.stream):
            try:
                with dist._coalescing_manager(group=self.group, device=self.flat.device, async_ops=False):
                    for i in indices:
                        dist.all_reduce(self.buckets[i], op=dist.ReduceOp.AVG, group=self.group)
            except (AttributeError, TypeError):
                for i in indices:
                    dist.all_reduce(self.buckets[i], op=dist.ReduceOp.AVG, group=self.group)
        self.launched.extend(indices)
        self.bytes_reduced += sum(self.buckets[i].numel() * 4 for i in indices)

    def wait(self):
        if self.stream is not None and self.launched:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.launched = []
