"""Image-parallel sharding of the hot path: one process per GPU, no data-path collective.

Every rank extracts the (replicated) prior shape and renders its own contiguous slice of the global batch - exactly
what the reference gets from accelerate/DDP (Trainer.py:170-180).  The only exchange is the all-reduce of parameter
gradients after the backward.  On one node that all-reduce is libb2a.so's own kernel over NVLink peer memory
(csrc/allreduce_p2p.cu: one launch per rank and bucket set; the ranks of this path are host-bound, so the host cost of issuing
the collective is what a step pays for it); NCCL through torch.distributed everywhere else.
"""
import ctypes
import os
import socket

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


class _DeviceArray:
    """A raw device pointer as a torch tensor (zero copy) through __cuda_array_interface__."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerMemory:
    """One buffer per rank that every rank of the node maps (CUDA IPC): [n floats | flag words], for b2a_allreduce_p2p.

    Construction is a collective (an all_gather_object of the 64-byte handles plus one agreement all-reduce): every rank of the
    group must construct it at the same point.  `ok` is False - on EVERY rank - when any rank could not set it up (ranks on
    different hosts, no peer access, more than 8 ranks, B2A_ALLREDUCE=nccl): callers then use NCCL."""
    CHANNELS = 8

    def __init__(self, n_floats, device, group=None):
        from . import _lib
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = (int(n_floats) + 3) // 4 * 4
        self.ptrs, self.own, self.ok = [None] * self.world, None, False
        self.channels = {}
        lib = _lib.lib()
        handle, err = b"", None
        try:
            if os.environ.get("B2A_ALLREDUCE", "p2p") == "nccl" or self.world > 8 or self.device.type != "cuda":
                raise RuntimeError("peer-memory all-reduce not selected")
            with torch.cuda.device(self.device):
                ptr, buf = ctypes.c_void_p(), ctypes.create_string_buffer(64)
                _lib.check(lib.b2a_p2p_alloc(4 * self.n + 4 * 16 * self.CHANNELS, ctypes.byref(ptr), buf))
                self.own, handle = ptr.value, bytes(buf.raw)
        except Exception as e:      # noqa: BLE001 - agreement below turns any local failure into a group-wide fallback
            err = e
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (socket.gethostname(), handle), group=group)
        good = err is None and all(h == gathered[0][0] and hd for h, hd in gathered)
        if good:
            try:
                with torch.cuda.device(self.device):
                    for p, (_, hd) in enumerate(gathered):
                        if p == self.rank:
                            self.ptrs[p] = self.own
                        else:
                            q = ctypes.c_void_p()
                            _lib.check(lib.b2a_p2p_open(hd, ctypes.byref(q)))
                            self.ptrs[p] = q.value
            except Exception:       # noqa: BLE001
                good = False
        agree = torch.tensor([1.0 if good else 0.0], device=self.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(agree.item() > 0.5)
        if self.ok:
            self.flat = torch.as_tensor(_DeviceArray(self.own, self.n), device=self.device)
            self.counters = torch.zeros(2 * self.CHANNELS, dtype=torch.int32, device=self.device)
            dist.barrier(group=group)
        else:
            self.close()

    def allreduce(self, start, stop, stream):
        """Average flat[start:stop] (start, stop multiples of 4) over the ranks, in place, one kernel on `stream`."""
        from . import _lib, ops
        key = (int(start), int(stop))
        ch = self.channels.get(key)
        if ch is None:
            if len(self.channels) >= self.CHANNELS:
                raise ValueError("PeerMemory: more than %d distinct ranges" % self.CHANNELS)
            i = len(self.channels)
            bufs = (ctypes.c_void_p * self.world)(*[p + 4 * key[0] for p in self.ptrs])
            flags = (ctypes.c_void_p * self.world)(*[p + 4 * self.n + 64 * i for p in self.ptrs])
            ch = self.channels[key] = [bufs, flags, self.counters.data_ptr() + 8 * i, 0]
        ch[3] += 1
        _lib.check(_lib.lib().b2a_allreduce_p2p(ch[0], ch[1], self.rank, self.world, key[1] - key[0], ch[3], ch[2], stream))
        ops.stats.count("b2a_allreduce_p2p", 1)

    def close(self):
        """Unmap the peers' buffers and free this rank's; raises if a collective gave up waiting for a peer (results undefined)."""
        from . import _lib
        lib = _lib.lib()
        failed = self.ok and self.own is not None and bool(self.counters[1::2].any().item())
        for p, q in enumerate(self.ptrs):
            if q is not None and p != self.rank:
                lib.b2a_p2p_close(q)
        if self.own is not None:
            lib.b2a_p2p_free(self.own)
        self.ptrs, self.own = [None] * self.world, None
        if failed:
            raise _lib.B2AError("b2a_allreduce_p2p: a peer rank did not arrive within the spin limit; reduced gradients are undefined")


class GradientBuckets:
    """DDP-style bucketed gradient all-reduce, overlapped with the backward pass (reference: accelerate / DistributedDataParallel,
    Trainer.py:170-180; SURVEY.md §2.2 sizes the MagicPony message at ~68 MB of fp32 parameter gradients per step).

    Flat fp32 buckets of at most `bucket_bytes`; `launch(i)` enqueues bucket i's all-reduce on a SIDE stream behind an
    event recorded on the compute stream, so the collective runs under whatever backward work is still queued; `wait()` makes the
    compute stream wait for every launched bucket - call it where the optimiser would read the gradients.  Outside a process
    group (N = 1) every method is a no-op.  The collective is libb2a.so's peer-memory kernel when the ranks share a node
    (`self.backend == "p2p"`: one launch for any contiguous run of buckets), NCCL AVG otherwise (`"nccl"`, c10d coalescing)."""

    def __init__(self, total_bytes, device, bucket_bytes=25 << 20, group=None, tail_bytes=0):
        """tail_bytes > 0: the LAST bucket holds exactly that many bytes (the gradients that become final last, e.g. d_sdf) and
        the rest of the set is cut into `bucket_bytes` buckets in front of it - the exposed collective is as small as possible."""
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        n = max(int(total_bytes) // 4, 1)
        per = max(int(bucket_bytes) // 4, 1)
        tail = min(max(int(tail_bytes) // 4, 0), n)
        head = (n - tail + 3) // 4 * 4 if tail else n       # the tail bucket starts 16-byte aligned (the peer-memory kernel moves float4)
        total = (head + tail + 3) // 4 * 4
        cuda = torch.device(device).type == "cuda"
        self.peer = PeerMemory(total, device, group) if (self.active and cuda) else None
        if self.peer is not None and not self.peer.ok:
            self.peer = None
        self.backend = "p2p" if self.peer is not None else ("nccl" if self.active else "none")
        self.storage = self.peer.flat if self.peer is not None else torch.zeros(total, device=device)
        self.flat = self.storage[:head + tail]
        self.bounds = [(i, min(i + per, n - tail)) for i in range(0, n - tail, per)] + ([(head, head + tail)] if tail else [])
        self.buckets = [self.flat[a:b] for a, b in self.bounds]
        self.stream = torch.cuda.Stream(device=device) if (self.active and cuda) else None
        self.ready = [torch.cuda.Event() for _ in self.buckets] if self.stream is not None else []
        self.launched = []
        self.bytes_reduced = 0

    def view(self, bucket, numel, offset=0):
        """A [numel] slice of a bucket: point a gradient at it so that the kernel writes straight into the bucket (no copy)."""
        return self.buckets[bucket][offset:offset + numel]

    def _peer_launch(self, indices, inline):
        """Contiguous runs of buckets -> one peer-memory kernel each; on the side stream behind the compute stream, or (`inline`)
        on the compute stream itself when nothing is left to overlap with."""
        runs = []
        for i in sorted(indices):
            a, b = self.bounds[i]
            b = (b + 3) // 4 * 4
            if runs and runs[-1][1] == a:
                runs[-1][1] = b
            else:
                runs.append([a, b])
        if inline:
            st = torch.cuda.current_stream()
        else:
            st = self.stream
            self.ready[indices[0]].record()
            st.wait_event(self.ready[indices[0]])
        for a, b in runs:
            self.peer.allreduce(a, b, st.cuda_stream)
        if not inline:
            self.launched.extend(indices)
        self.bytes_reduced += sum(self.buckets[i].numel() * 4 for i in indices)

    def launch(self, i):
        if not self.active:
            return
        if self.peer is not None:
            return self._peer_launch((i,), False)
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

    def launch_many(self, indices, inline=False):
        """Buckets `indices` in ONE enqueue.  p2p: one kernel per contiguous run of buckets (`inline`: on the compute stream, for a
        set with nothing left to overlap - saves the event and the two stream waits).  nccl: the collectives are issued inside one
        NCCL group (c10d coalescing manager), i.e. one launch on the side stream instead of one c10d call + launch per bucket.
        (Capturing the NCCL collectives into a CUDA graph was tried and deadlocked on the 2-GPU box; not used.)"""
        indices = tuple(indices)
        if not self.active or not indices:
            return
        if self.peer is not None:
            return self._peer_launch(indices, inline)
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
