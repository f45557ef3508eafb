"""N > 1 host logic on CPU (gloo, world_size 2): the image-parallel shard split and the gradient all-reduce that is the
path's only exchange.  The kernels themselves never communicate."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import pkg


def test_shard_range_partitions_the_batch():
    par = pkg("parallel")
    for n in (1, 7, 16, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib
        par = importlib.import_module("3danimals_b200.parallel")
        # per-rank stand-ins for d_sdf / d_articulation of this rank's image shard
        lo, hi = par.shard_range(6, rank, world)
        rng = np.random.RandomState(0)
        per_image = rng.randn(6, 50).astype(np.float32)            # same on every rank
        d_sdf = torch.from_numpy(per_image[lo:hi].sum(0).copy())
        d_bias = torch.full((3,), float(rank + 1))
        nbytes = par.allreduce_gradients([d_sdf, None, d_bias], average=True)
        assert nbytes == (50 + 3) * 4
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([d_sdf.numpy(), d_bias.numpy()]))
        # bucketed form (bench.py's DDP stand-in): 3 buckets of <= 40 bytes over a 100-byte gradient set, launched one by one
        gb = par.GradientBuckets(100, "cpu", bucket_bytes=40)
        assert gb.active and [b.numel() for b in gb.buckets] == [10, 10, 5]
        gb.view(2, 5).copy_(torch.arange(5.0) * (rank + 1))
        gb.buckets[0].fill_(float(rank))
        for i in (0, 2):
            gb.launch(i)
        gb.wait()
        assert gb.bytes_reduced == 60 and gb.launched == []
        np.save(os.path.join(out_dir, "b%d.npy" % rank), gb.flat.numpy())
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)                                   # every rank holds the same reduced gradient
    per_image = np.random.RandomState(0).randn(6, 50).astype(np.float32)
    assert np.allclose(r0[:50], per_image.sum(0) / world, atol=1e-6)   # = DDP mean over ranks of the per-shard sums
    assert np.allclose(r0[50:], 1.5)
    b0, b1 = np.load(tmp_path / "b0.npy"), np.load(tmp_path / "b1.npy")
    assert np.array_equal(b0, b1) and np.allclose(b0[:10], 0.5) and np.allclose(b0[10:20], 0) and np.allclose(b0[20:], np.arange(5.0) * 1.5)


def test_allreduce_is_noop_without_group():
    par = pkg("parallel")
    g = torch.ones(4)
    assert par.allreduce_gradients([g]) == 0 and torch.equal(g, torch.ones(4))
    gb = par.GradientBuckets(64, "cpu", bucket_bytes=32)
    gb.launch(0); gb.wait()
    assert not gb.active and gb.bytes_reduced == 0 and len(gb.buckets) == 2
