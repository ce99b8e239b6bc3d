"""Per-epoch LR schedule wiring — mirrors /root/reference/optimizers/scheduler.py:4-62 and
main.build_lr_schedule (main.py:279-300): a linear warm-up (from 0, so the whole first epoch trains with lr = 0 —
SURVEY.md Q11) that hands over to the wrapped schedule (cosine over epochs - warmup, or fixed).  Host-side only;
it changes `param_groups[i]['lr']`, which `byol_b200.lars.LARS.step` uploads to the fused kernel when it changes.
"""
from torch.optim import lr_scheduler
from torch.optim.lr_scheduler import LambdaLR


class LinearWarmup(LambdaLR):
    """Linearly increases the LR factor from 0 to 1 over `warmup_steps` scheduler steps, then stays at 1."""

    def __init__(self, optimizer, warmup_steps, last_epoch=-1):
        self.warmup_steps = warmup_steps
        self.complete = False
        super(LinearWarmup, self).__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1.0, self.warmup_steps))
        self.complete = True
        return 1.


class Scheduler(object):
    """Container: steps the warm-up until it is complete, then the normal scheduler."""

    def __init__(self, normal_schededuler, warmup_scheduler=None):
        self.warmup = warmup_scheduler
        self.sched = normal_schededuler

    def get_last_lr(self):
        if self.warmup is not None and not self.warmup.complete:
            return self.warmup.get_last_lr()
        return self.sched.get_last_lr()

    def state_dict(self):
        return {'warmup': self.warmup.state_dict() if self.warmup is not None else {},
                'sched': self.sched.state_dict()}

    def load_state_dict(self, state_dict):
        if self.warmup:
            self.warmup.load_state_dict(state_dict['warmup'])
        self.sched.load_state_dict(state_dict['sched'])

    def step(self, *args, **kwargs):
        if self.warmup is not None and not self.warmup.complete:
            return self.warmup.step(*args, **kwargs)
        return self.sched.step(*args, **kwargs)


def build_lr_schedule(optimizer, epochs, warmup=10, lr_update_schedule='cosine', last_epoch=-1):
    """main.py:279-300."""
    if lr_update_schedule == 'fixed':
        sched = lr_scheduler.LambdaLR(optimizer, lambda epoch: 1.0, last_epoch=last_epoch)
    elif lr_update_schedule == 'cosine':
        sched = lr_scheduler.CosineAnnealingLR(optimizer, T_max=epochs - warmup, last_epoch=last_epoch)
    else:
        raise NotImplementedError("lr scheduler {} not implemented".format(lr_update_schedule))
    if warmup > 0:
        sched = Scheduler(sched, LinearWarmup(optimizer, warmup_steps=warmup, last_epoch=last_epoch))
    return sched
