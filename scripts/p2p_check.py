"""Peer-memory all-reduce against NCCL, N ranks of one node (torchrun).  Correctness over many rounds and sizes, then host and device cost
of both backends for the two message sizes of the bench (stand-in set, d_sdf tail).  Run under `timeout`."""
import importlib
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
par = importlib.import_module("3danimals_b200.parallel")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tail = 2146689
    gb = par.GradientBuckets(68 << 20, dev, tail_bytes=4 * tail)
    if rank == 0:
        print("backend", gb.backend, "buckets", [b.numel() for b in gb.buckets], flush=True)
    assert gb.backend == "p2p", gb.backend
    last = len(gb.buckets) - 1
    g = torch.Generator(device=dev)
    worst = 0.0
    rounds = int(os.environ.get("P2P_CHECK_ROUNDS", "200"))
    for it in range(rounds):
        g.manual_seed(1000 * it + rank)
        src = torch.randn(gb.flat.numel(), device=dev, generator=g)
        gb.flat.copy_(src)
        want = src.clone()
        dist.all_reduce(want, op=dist.ReduceOp.AVG)
        if it % 2 == 0:
            gb.launch_many(range(last))
            gb.launch_many((last,), inline=True)
        else:
            gb.launch_many(range(last + 1))
        gb.wait()
        for i, (a, b) in enumerate(gb.bounds):
            worst = max(worst, float((gb.flat[a:b] - want[a:b]).abs().max()))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank holds bit-identical results
    chk = gb.flat.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("max |p2p - nccl| over %d rounds: %.3e   identical across ranks: %s" % (rounds, float(t), bool(lo == hi)), flush=True)
    assert float(t) < 1e-6 and bool(lo == hi)

    if os.environ.get("P2P_CHECK_TIMING", "1") == "0":
        dist.barrier()
        gb.peer.close()
        dist.destroy_process_group()
        return

    def timed(fn, n=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); host = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        return host, e0.elapsed_time(e1) / n * 1e3

    nccl = par.GradientBuckets.__new__(par.GradientBuckets)
    nccl.__dict__.update(gb.__dict__)
    nccl.peer, nccl.backend = None, "nccl"
    nccl.flat = torch.zeros_like(gb.flat)
    nccl.buckets = [nccl.flat[a:b] for a, b in gb.bounds]
    rows = []
    for name, obj in (("p2p", gb), ("nccl", nccl)):
        def heads(o=obj):
            o.launch_many(range(last)); o.wait()
        def tails(o=obj):
            o.launch_many((last,), inline=True); o.wait()
        rows.append((name, "stand-in set %.1f MB" % (sum(b.numel() for b in obj.buckets[:last]) * 4 / 1e6),) + timed(heads))
        rows.append((name, "tail %.1f MB" % (obj.buckets[last].numel() * 4 / 1e6),) + timed(tails))
    if rank == 0:
        for r in rows:
            print("%-5s %-24s host %.1f us/call   device %.1f us/call" % r, flush=True)
    dist.barrier()
    gb.peer.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
