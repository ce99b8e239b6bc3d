"""CPU, world_size 2 over gloo: the two data-parallel exchanges of the hot path (byol_b200/comm.py) and the
oracle's multi-rank emulation they are checked against."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from byol_b200 import comm
    assert comm.world_size() == world and comm.rank() == rank
    # (1) SyncBN statistics: SUM of per-rank [sum, sqsum] for several lanes in one call
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(16, 8, generator=g)
    stats = torch.cat([x.sum(0), (x * x).sum(0)])
    local = torch.empty_like(stats)
    comm.allreduce_sum_(stats, local_out=local, channel=1)      # CPU tensors: the gloo path; local_out = the input
    assert torch.equal(local, torch.cat([x.sum(0), (x * x).sum(0)]))
    assert comm.peer_exchange(torch.device("cpu")) is None       # the NVLink peer exchange needs NCCL + CUDA
    # (2) DDP: MEAN of the flat gradient
    grad = torch.full((1000,), float(rank + 1))
    comm.allreduce_mean_(grad)
    ret[rank] = (stats.clone(), grad.clone(), x.clone())
    dist.destroy_process_group()


def test_two_rank_exchanges_gloo():
    world, port = 2, 29533
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    s0, g0, x0 = ret[0]
    s1, g1, x1 = ret[1]
    assert torch.equal(s0, s1) and torch.equal(g0, g1)           # replicas stay identical
    full = torch.cat([x0, x1])
    assert torch.allclose(s0[:8], full.sum(0), atol=1e-5)
    assert torch.allclose(s0[8:], (full * full).sum(0), atol=1e-4)
    assert torch.allclose(g0, torch.full((1000,), 1.5))          # mean of 1 and 2


def test_oracle_multirank_emulation_matches_per_rank_losses():
    """Q2: under data parallelism the loss norms are over the RANK-LOCAL shard; the oracle's world>1 emulation
    (global BN statistics, per-shard losses, averaged) must differ from treating the global batch as one shard."""
    from oracle import byol_oracle as O
    g = torch.Generator().manual_seed(0)
    q1, q2, z1, z2 = [torch.randn(8, 16, generator=g) for _ in range(4)]
    per_rank = torch.stack([O.loss_function(q1[r * 4:(r + 1) * 4], q2[r * 4:(r + 1) * 4], z1[r * 4:(r + 1) * 4],
                                            z2[r * 4:(r + 1) * 4]) for r in range(2)]).mean()
    whole = O.loss_function(q1, q2, z1, z2)
    assert abs(per_rank.item() - whole.item()) > 1e-4
