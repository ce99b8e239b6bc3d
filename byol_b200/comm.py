"""Data-parallel plumbing (one process per GPU, torch.distributed): the only two exchanges the hot path has.

* SyncBatchNorm statistics (main.py:433 -> torch/nn/modules/_functions.py:49-74,158-159): SUM all-reduce of the
  per-channel partial sums of all lanes of one layer in ONE call.
* DDP gradient averaging (main.py:440-443, 617): ONE mean all-reduce of the flat gradient buffer per backward.

NCCL over NVLink/NVSwitch on the GPUs; the same functions run on gloo/CPU tensors for the world_size-2 tests.
"""
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def allreduce_sum_(t):
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_mean_(t):
    w = world_size()
    if w > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:   # gloo has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(w)
    return t
