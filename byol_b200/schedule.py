"""Per-epoch learning-rate schedule in closed form.

What the reference wires up in main.build_lr_schedule (/root/reference/main.py:279-300, stepped once per epoch at
main.py:763) is, as a function of the number of completed epochs e: a linear ramp base*e/W for e <= W (so the whole
first epoch trains with lr = 0) followed by either a constant or half a cosine period over the
remaining E - W epochs.  Here that function is evaluated directly — one stateless formula plus a counter — instead
of chaining torch.optim.lr_scheduler objects; tests/test_host_logic.py checks it against the learning rates the
reference's own wiring produced (tests/golden/lr_schedule.npz).

`EpochSchedule.step()` only rewrites `param_groups[i]['lr']` on the host; `byol_b200.lars.LARS.step` uploads the
per-tensor rates to the fused kernel whenever they change.
"""
import math


def lr_factor(epoch, epochs, warmup=10, kind="cosine"):
    """Multiplier of the base LR after `epoch` completed epochs."""
    if kind not in ("cosine", "fixed"):
        raise NotImplementedError("lr scheduler %s not implemented" % kind)
    if warmup > 0 and epoch < warmup:
        return epoch / float(warmup)
    if kind == "fixed":
        return 1.0
    span = max(1, epochs - warmup)
    return 0.5 * (1.0 + math.cos(math.pi * (epoch - max(warmup, 0)) / span))


class EpochSchedule(object):
    """Holds the base LR of every param group and the epoch counter; accepts a plain optimizer or byol_b200.LARS."""

    def __init__(self, optimizer, epochs, warmup=10, kind="cosine", last_epoch=0):
        self.optimizer, self.epochs, self.warmup, self.kind = optimizer, epochs, warmup, kind
        self.base_lrs = [g["lr"] for g in optimizer.param_groups]
        self.epoch = last_epoch
        self._write()

    def _write(self):
        f = lr_factor(self.epoch, self.epochs, self.warmup, self.kind)
        for g, base in zip(self.optimizer.param_groups, self.base_lrs):
            g["lr"] = base * f

    def step(self):
        self.epoch += 1
        self._write()

    def get_last_lr(self):
        return [g["lr"] for g in self.optimizer.param_groups]

    def state_dict(self):
        return {"epoch": self.epoch, "base_lrs": list(self.base_lrs)}

    def load_state_dict(self, state):
        self.epoch, self.base_lrs = state["epoch"], list(state["base_lrs"])
        self._write()
