"""Shims for helpers.layers (call sites: /root/reference/main.py:321,436,440,499,586,626,749-753)."""
import torch
import torch.nn as nn


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    """Two param groups; biases / 1-d (BN) params get no decay and are flagged 'ignore' for LARS
    (/root/reference/optimizers/lars.py:88,99-100).  Behaviour inferred — the submodule source is absent."""
    decay, no_decay = [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith(".bias") or name in skip_list:
            no_decay.append(param)
        else:
            decay.append(param)
    return [
        {"params": no_decay, "weight_decay": 0.0, "ignore": True},
        {"params": decay, "weight_decay": weight_decay, "ignore": False},
    ]


def init_weights(module, init=None):          # main.py:436 (None -> keep default torch init)
    return module


class DistributedDataParallelPassthrough(nn.parallel.DistributedDataParallel):   # main.py:440
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)


def append_save_and_load_fns(model, optimizer, scheduler, grapher, args):        # main.py:749
    return model


class ModelSaver(object):                                                        # main.py:750-753
    def __init__(self, model, early_stop=False, rank=0, burn_in_interval=0, larger_is_better=False,
                 max_early_stop_steps=10, **kwargs):
        self.model = model

    def restore(self):
        return {"epoch": 1}

    def __call__(self, loss):
        return False


def polyak_ema_parameters(model, decay):                                         # main.py:499,626 (off by default)
    return None


def get_polyak_prediction(model, pred_fn):                                       # main.py:586 (off by default)
    return pred_fn()
