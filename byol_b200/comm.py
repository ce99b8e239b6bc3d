"""Data-parallel plumbing (one process per GPU, torch.distributed): the only two exchanges the hot path has.

* SyncBatchNorm statistics (main.py:433 -> torch/nn/modules/_functions.py:49-74,158-159): SUM all-reduce of the
  per-channel partial sums of all lanes of one layer in ONE call.
* DDP gradient averaging (main.py:440-443, 617): ONE mean all-reduce of the flat gradient buffer per backward.

NCCL over NVLink/NVSwitch on the GPUs; the same functions run on gloo/CPU tensors for the world_size-2 tests.
"""
import os

import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


class PeerExchange(object):
    """Rank-ordered sum of small per-layer statistic vectors over NVLink peer memory: one single-CTA kernel per
    exchange (csrc/xchg.cu) instead of an NCCL launch — deterministic, identical bits on every rank, and capturable
    in a CUDA graph (the NCCL version of this exchange is ~110 latency-bound collectives per step)."""

    def __init__(self, device, cap_bytes):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import lib
        slots, maxw, flag_bytes = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib.byol_xchg_layout(ctypes.byref(slots), ctypes.byref(maxw), ctypes.byref(flag_bytes))
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > maxw.value:
            raise RuntimeError("peer exchange supports up to %d ranks" % maxw.value)
        self.cap_bytes = (int(cap_bytes) + 1023) // 1024 * 1024
        nbytes = flag_bytes.value + slots.value * self.cap_bytes
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()          # every rank has zeroed its flags before anyone signals

    def sum_(self, t, local_out=None):
        from ._lib import lib, check
        if t.dtype not in (torch.float32, torch.float64) or not t.is_contiguous():
            raise ValueError("PeerExchange.sum_: contiguous fp32 / fp64 tensors only")
        check(lib.byol_xchg_sum(t.data_ptr(), 0 if local_out is None else local_out.data_ptr(), t.numel(),
                                int(t.dtype == torch.float64), self.ptrs, self.world, self.rank, self.cap_bytes,
                                self.counter.data_ptr(), torch.cuda.current_stream().cuda_stream), "byol_xchg_sum")
        return t


NUM_CHANNELS = 2      # independent exchange sequences: one per CUDA stream that carries SyncBatchNorm layers
_peer = {"tried": False, "xchg": None}


def peer_exchange(device=None, cap_bytes=1 << 19):
    """The process-wide list of PeerExchange channels (created on first use, collectively: every rank must call it at
    the same point), or None when disabled (BYOL_B200_PEER_XCHG=0), unavailable (no NCCL / symmetric memory) or
    world == 1.  Each channel has its own buffer and sequence counter, so two streams (the online / target lane pairs
    of the forward pass, the two views of the backward pass) can exchange concurrently."""
    if _peer["tried"]:
        return _peer["xchg"]
    _peer["tried"] = True
    if world_size() <= 1 or dist.get_backend() != "nccl" or os.environ.get("BYOL_B200_PEER_XCHG", "1") == "0":
        return None
    ok = torch.ones(1, device=device)
    try:
        x = [PeerExchange(device, cap_bytes) for _ in range(NUM_CHANNELS)]
    except Exception as e:      # symmetric memory needs P2P / fabric handles; fall back to NCCL on ALL ranks
        print("[byol_b200] peer exchange unavailable (%s: %s); SyncBatchNorm statistics use NCCL" % (type(e).__name__, e))
        x = None
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    _peer["xchg"] = x if bool(ok.item() > 0) else None
    return _peer["xchg"]


def allreduce_sum_(t, local_out=None, channel=0):
    """In-place SUM over ranks of a small statistics vector; local_out (optional) receives this rank's input.
    `channel` selects the exchange sequence (all ranks must use the same channel for the same exchange)."""
    if world_size() > 1:
        x = peer_exchange(t.device) if t.is_cuda else None
        if x is not None and t.numel() * t.element_size() <= x[channel].cap_bytes:
            return x[channel].sum_(t, local_out)
        if local_out is not None:
            local_out.copy_(t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    elif local_out is not None:
        local_out.copy_(t)
    return t


def uses_nccl_for_statistics(device):
    """True if the per-layer exchanges go through NCCL (then the step is not captured in a CUDA graph)."""
    return world_size() > 1 and peer_exchange(device) is None


def allreduce_mean_(t):
    w = world_size()
    if w > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:   # gloo has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(w)
    return t
