"""The peer-memory gradient all-reduce (csrc/allreduce_p2p.cu, parallel.PeerMemory / GradientBuckets).  Stands where the reference has
DistributedDataParallel's NCCL all-reduce (Trainer.py:170-180), so NCCL AVG on the same data is the checker.  The 2-rank case needs
two GPUs and is skipped on a single-GPU box (scripts/p2p_check.py is the same check, run at N = 2 and N = 8 in profiles/)."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist

from conftest import pkg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_peer_memory_single_rank(cuda):
    """world = 1: the whole protocol (allocation, IPC handle, flag words, both barriers, the grid barrier) with itself as only peer."""
    par, ops = pkg("parallel"), pkg("ops")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % _free_port(), rank=0, world_size=1)
    try:
        pm = par.PeerMemory(100003, cuda)
        assert pm.ok and pm.flat.numel() == 100004 and float(pm.flat.abs().max()) == 0.0
        src = torch.randn(100004, device=cuda)
        pm.flat.copy_(src)
        ops.stats.reset()
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            pm.allreduce(0, 100004, st)         # average over one rank: the identity, three epochs of the same channel
        pm.allreduce(4000, 8000, st)            # a second channel on a sub-range
        torch.cuda.synchronize()
        assert torch.equal(pm.flat, src) and ops.stats.calls["b2a_allreduce_p2p"] == 4 and len(pm.channels) == 2
        with pytest.raises(Exception):
            pm.allreduce(2, 10, st)             # not 16-byte aligned
        pm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_two_ranks_matches_nccl():
    env = dict(os.environ, P2P_CHECK_ROUNDS="20", P2P_CHECK_TIMING="0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "p2p_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "identical across ranks: True" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
