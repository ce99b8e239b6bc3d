"""GPU, 2 ranks over NCCL (skipped with < 2 GPUs): SyncBatchNorm statistic exchange + flat gradient all-reduce.

Each rank runs the CUDA training step on its shard; checks: (1) parameters, EMA target and BN buffers stay
BIT-identical across ranks; (2) outputs / loss agree with the oracle's multi-rank emulation (global BN statistics,
rank-local loss norms (Q2), rank-averaged gradients) at the bf16 tolerance of tests/test_gpu_step.py.
"""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ARCH, REP, B, R, SEED, LR = "resnet:bottleneck:2,1,1,1", 2048, 8, 64, 41, 0.3


def _data(world):
    g = torch.Generator().manual_seed(77)
    a1 = torch.rand(world * B, 3, R, R, generator=g)
    a2 = torch.rand(world * B, 3, R, R, generator=g)
    lab = torch.randint(0, 1000, (world * B,), generator=g)
    return a1, a2, lab


STEPS = 3     # step 1 eager, step 2 captured into CUDA graphs and replayed, step 3 replayed


def _worker(rank, world, port, ret, peer_xchg):
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)      # a cross-rank deadlock must not eat the GPU lease
    os.environ["BYOL_B200_PEER_XCHG"] = "1" if peer_xchg else "0"
    import torch.distributed as dist
    import torch.nn as nn
    import torch.nn.functional as F
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from byol_b200.model import BYOL
    from byol_b200 import wiring
    torch.manual_seed(SEED)
    model = BYOL(REP, 256, 1000, 10, arch=ARCH)
    model = nn.SyncBatchNorm.convert_sync_batchnorm(model).cuda().train()
    net = wiring.DistributedDataParallelPassthrough(model)
    opt = wiring.LARS(torch.optim.SGD(wiring.add_weight_decay(model, 1e-6), lr=LR, momentum=0.9), eps=0.0)
    a1, a2, lab = _data(world)
    sl = slice(rank * B, (rank + 1) * B)
    out = None
    for _ in range(STEPS):
        stats = wiring.train_step(net, opt, a1[sl].cuda(), a2[sl].cuda(), lab[sl].cuda())
    torch.cuda.synchronize()
    from byol_b200 import comm
    captured = any(v != "warm" for v in model._engine.graphs.values())
    assert (comm.peer_exchange(torch.device("cuda", rank)) is not None) == bool(peer_xchg), "exchange path mismatch"
    assert captured == bool(peer_xchg), "graphs are used exactly when the statistics exchange is capturable"
    sd = model.state_dict()
    ret[rank] = {"theta": model._engine.theta.cpu(), "ema": model.target_network.mean.cpu(),
                 "rm": sd["base_network.1.running_mean"].cpu(), "loss": float(stats["loss_mean"]),
                 "byol": float(stats["byol_loss_mean"])}
    dist.destroy_process_group()


@pytest.mark.parametrize("peer_xchg", [False, True], ids=["nccl-stats-eager", "peer-exchange-graphs"])
def test_two_rank_syncbn_ddp(cuda, peer_xchg):
    """SyncBatchNorm statistics over (a) per-layer NCCL all-reduces, eager launches; (b) the peer-memory exchange
    kernel (csrc/xchg.cu) inside CUDA-graph replays.  Both must keep the replicas bit-identical and follow the oracle's
    2-rank emulation."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29544 + int(peer_xchg), ret, peer_xchg), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert torch.equal(r0["theta"], r1["theta"]), "replicas diverged (parameters)"
    assert torch.equal(r0["ema"], r1["ema"]), "replicas diverged (EMA target)"
    assert torch.equal(r0["rm"], r1["rm"]), "replicas diverged (BN running mean)"
    # oracle emulation of the same 2-rank job
    from oracle import byol_oracle as O
    params, buffers = O.init_reference_state(ARCH, SEED)
    theta0 = torch.cat([p.reshape(-1) for p in params.values()])
    oracle = O.OracleBYOL(ARCH, params, buffers, 10, storage="bf16")
    a1, a2, lab = _data(world)
    for _ in range(STEPS):
        ref = oracle.train_step(a1, a2, lab, LR, world=world, sync_bn=True)
    upd, upd_ref = r0["theta"] - theta0, oracle.flat_params() - theta0
    cos = float((upd.double() @ upd_ref.double()) / (upd.double().norm() * upd_ref.double().norm()))
    mean_loss = 0.5 * (r0["loss"] + r1["loss"])
    print("2-rank: loss %.5f (oracle %.5f)  update cosine %.5f  rank byol losses %.6f %.6f" %
          (mean_loss, float(ref["loss"]), cos, r0["byol"], r1["byol"]))
    assert abs(mean_loss - float(ref["loss"])) < 2e-2 * abs(float(ref["loss"]))
    assert cos > 0.9
    assert r0["byol"] != r1["byol"]        # Q2: the loss (and its norms) are rank-local
    e = float((r0["rm"] - oracle.buffers["base_network.1.running_mean"]).abs().max() /
              oracle.buffers["base_network.1.running_mean"].abs().max())
    assert e < 2e-2, e                      # global (cross-rank) BN statistics
