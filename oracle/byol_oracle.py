"""CPU oracle: a functional fp32 restatement of the reference's BYOL training step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — the product (byol_b200/) never imports this.

Parity status: PINNED.  The reference repository has no tests or golden vectors of its own, so
the pins are outputs of the unmodified reference (/root/reference/main.py, objective.py, optimizers/lars.py)
executed in the build container by tests/golden/make_golden.py and committed as tests/golden/*.npz;
tests/test_oracle_golden.py checks this restatement against them.  One boundary stays UNPINNED: the
param-group split of the missing `helpers.layers.add_weight_decay` submodule is inferred
from /root/reference/optimizers/lars.py:88,99-100 ("ignore bias / bn terms").

What is restated, with the reference lines each function follows:

* parameter creation order / init      main.py:190-212  (torchvision resnet children[:-1], head, predictor,
                                                          linear_classifier; flat order = registration order)
* encoder / MLP forward                main.py:229-240  + torchvision/models/resnet.py (v1.5 topology)
* 4-pass forward, classifier, EMA      main.py:242-276, main.py:214-227 (target = same graph at EMA weights)
* CosEMA                               main.py:133-164  (float64 cosine decay, three fp32 ops, step counter)
* loss                                 objective.py:6-25 (Frobenius-normalised, rank-local)
* step loop                            main.py:579-631  (byol + CE loss, zero_grad, backward, optimizer.step)
* LARS + SGD momentum                  optimizers/lars.py:84-127, main.py:316,332-340
* SyncBatchNorm / DDP emulation        main.py:433,440 with equal per-rank counts: BN over the global per-view
                                       batch, losses (and their Frobenius norms) per rank shard, gradients
                                       averaged over ranks
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ARCHS = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
    "resnet152": ("bottleneck", [3, 8, 36, 3]),
    "resnet200": ("bottleneck", [3, 24, 36, 3]),   # BASELINE.json config 5; not a torchvision name
}


def arch_spec(arch):
    """(block kind, stage depths).  Besides the named nets, "resnet:<basic|bottleneck>:d1,d2,d3,d4" describes a
    custom-depth ResNet (used for shallow, well-conditioned parity tests)."""
    if arch in ARCHS:
        return ARCHS[arch]
    _, kind, depths = arch.split(":")
    return kind, [int(d) for d in depths.split(",")]


def representation_size(arch):
    return 512 if arch_spec(arch)[0] == "basic" else 2048


def build_reference_modules(arch, representation, projection=256, head_latent=4096, num_classes=1000):
    """Instantiate the same nn.Modules in the same order as main.py:190-208 (so that, under a fixed
    torch.manual_seed, the initial parameters equal the reference's).  Returns an nn.Module whose
    named_parameters()/named_buffers() follow the reference's registration order."""
    import torchvision
    kind, depths = arch_spec(arch)
    if arch in torchvision.models.__dict__:
        net = torchvision.models.__dict__[arch](weights=None)
    else:
        from torchvision.models.resnet import ResNet, Bottleneck, BasicBlock
        net = ResNet(Bottleneck if kind == "bottleneck" else BasicBlock, depths)

    class _Container(nn.Module):
        pass

    m = _Container()
    m.base_network = nn.Sequential(*list(net.children())[:-1])
    m.head = nn.Sequential(nn.Linear(representation, head_latent), nn.BatchNorm1d(head_latent), nn.ReLU(),
                           nn.Linear(head_latent, projection))
    m.predictor = nn.Sequential(nn.Linear(projection, head_latent), nn.BatchNorm1d(head_latent), nn.ReLU(),
                                nn.Linear(head_latent, projection))
    m.linear_classifier = nn.Linear(representation, num_classes)
    return m


def init_reference_state(arch, seed, projection=256, head_latent=4096, num_classes=1000):
    """(params OrderedDict name->fp32 tensor in flat order, buffers OrderedDict) as the reference creates them."""
    torch.manual_seed(seed)
    m = build_reference_modules(arch, representation_size(arch), projection, head_latent, num_classes)
    params = OrderedDict((k, v.detach().clone()) for k, v in m.named_parameters())
    buffers = OrderedDict((k, v.detach().clone()) for k, v in m.named_buffers())
    return params, buffers


def is_ignored(name, tensor):
    """helpers.layers.add_weight_decay (inferred): 1-d tensors and biases get no decay / no LARS scaling."""
    return tensor.dim() == 1 or name.endswith(".bias")


def cos_ema_decay(step, total_steps, base_decay):
    """main.py:159 — evaluated in numpy float64."""
    return 1 - (1 - base_decay) * (np.cos(np.pi * step / total_steps) + 1) / 2.0


def regression_loss(x, y):
    """objective.py:6-9: norms are over the WHOLE matrix (Q1)."""
    return -2 * torch.sum(x * y, dim=-1) / (x.norm() * y.norm())


def loss_function(online_prediction1, online_prediction2, target_projection1, target_projection2):
    """objective.py:12-25."""
    loss_ab = regression_loss(online_prediction1, target_projection2.detach())
    loss_ba = regression_loss(online_prediction2, target_projection1.detach())
    return torch.mean(loss_ab + loss_ba)


def _identity(t):
    return t


def bf16_storage(t):
    """Round to bf16 with a straight-through gradient: emulates the CUDA path's bf16 activation / weight STORAGE
    (fp32 accumulation everywhere), so that parity against the kernels can be asserted tightly.  The fp32 oracle
    (storage='fp32') stays the reference-faithful one that is pinned against the golden vectors."""
    return t + (t.detach().to(torch.bfloat16).to(torch.float32) - t.detach())


class _BN(object):
    """torch.nn.functional.batch_norm with nn.BatchNorm's bookkeeping (momentum 0.1, eps 1e-5, unbiased running
    variance, num_batches_tracked += 1 per training call).  Under emulated SyncBatchNorm the caller passes the
    global per-view batch, which is what SyncBatchNorm's count-weighted statistics equal for equal shards."""

    def __init__(self, buffers, eps=1e-5, momentum=0.1):
        self.buffers, self.eps, self.momentum = buffers, eps, momentum

    def __call__(self, x, prefix, P, train):
        rm, rv = self.buffers[prefix + ".running_mean"], self.buffers[prefix + ".running_var"]
        if train:
            self.buffers[prefix + ".num_batches_tracked"] += 1
        return F.batch_norm(x, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], train, self.momentum, self.eps)


def block_list(arch, prefix="base_network"):
    """[(parameter prefix, stride)] of the residual blocks in execution order."""
    kind, depths = arch_spec(arch)
    return [("%s.%d.%d" % (prefix, 4 + li, bi), 2 if (li > 0 and bi == 0) else 1)
            for li, depth in enumerate(depths) for bi in range(depth)]


def stem_forward(P, bn, x, train, prefix="base_network", q=_identity, trace=None):
    """conv 7x7/2 -> BN -> ReLU -> maxpool 3x3/2 (torchvision resnet stem; children 0-3 of main.py:190-193)."""
    x = q(F.conv2d(q(x), q(P[prefix + ".0.weight"]), None, 2, 3))
    if trace is not None:
        trace.append(("stem_conv", x.detach()))
    x = q(torch.relu(bn(x, prefix + ".1", P, train)))
    if trace is not None:
        trace.append(("stem_act", x.detach()))
    return F.max_pool2d(x, 3, 2, 1)


def block_forward(kind, P, bn, x, p, stride, train, q=_identity):
    """One torchvision BasicBlock / Bottleneck (v1.5: the stride sits on the 3x3)."""
    conv = lambda inp, name, st=1, pad=0: q(F.conv2d(inp, q(P[name]), None, st, pad))
    identity = x
    if kind == "bottleneck":
        out = q(torch.relu(bn(conv(x, p + ".conv1.weight"), p + ".bn1", P, train)))
        out = q(torch.relu(bn(conv(out, p + ".conv2.weight", stride, 1), p + ".bn2", P, train)))
        out = bn(conv(out, p + ".conv3.weight"), p + ".bn3", P, train)
    else:
        out = q(torch.relu(bn(conv(x, p + ".conv1.weight", stride, 1), p + ".bn1", P, train)))
        out = bn(conv(out, p + ".conv2.weight", 1, 1), p + ".bn2", P, train)
    if (p + ".downsample.0.weight") in P:
        identity = bn(conv(x, p + ".downsample.0.weight", stride), p + ".downsample.1", P, train)
    return q(torch.relu(out + identity))


def encoder_forward(arch, P, bn, x, train, prefix="base_network", trace=None, q=_identity):
    """torchvision ResNet children[:-1] as an nn.Sequential: indices 0 conv1, 1 bn1, 2 relu, 3 maxpool,
    4-7 layer1-4, 8 avgpool (main.py:190-193, 237)."""
    kind, _ = arch_spec(arch)
    x = stem_forward(P, bn, x, train, prefix, q, trace)
    if trace is not None:
        trace.append(("pool", x.detach()))
    for p, stride in block_list(arch, prefix):
        x = block_forward(kind, P, bn, x, p, stride, train, q)
        if trace is not None:
            trace.append((p, x.detach()))
    return x.mean((2, 3))  # AdaptiveAvgPool2d(1) + view(-1, C)  (main.py:237)


def mlp_forward(P, bn, x, prefix, train, q=_identity):
    """main.py:194-205: Linear -> BatchNorm1d -> ReLU -> Linear."""
    h = q(F.linear(q(x), q(P[prefix + ".0.weight"]), P[prefix + ".0.bias"]))
    h = q(torch.relu(bn(h, prefix + ".1", P, train)))
    return F.linear(h, q(P[prefix + ".3.weight"]), P[prefix + ".3.bias"])


def prediction(arch, P, bn, aug, train, q=_identity):
    """main.py:229-240."""
    representation = encoder_forward(arch, P, bn, aug, train, q=q)
    projection = mlp_forward(P, bn, representation, "head", train, q=q)
    pred = mlp_forward(P, bn, projection, "predictor", train, q=q)
    return representation, projection, pred


class OracleBYOL(object):
    """State + one-step semantics of main.BYOL + LARS(SGD) for `world` emulated data-parallel ranks."""

    def __init__(self, arch, params, buffers, total_training_steps, base_decay=0.996, weight_decay=1e-6,
                 momentum=0.9, trust_coef=0.001, lars_eps=0.0, storage="fp32"):
        self.arch = arch
        assert storage in ("fp32", "bf16")
        self.q = bf16_storage if storage == "bf16" else _identity
        self.names = list(params.keys())
        self.params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())
        self.buffers = OrderedDict((k, v.detach().clone()) for k, v in buffers.items())
        self.bn = _BN(self.buffers)
        self.total_steps = total_training_steps
        self.base_decay = base_decay
        # main.py:211-212: the constructor calls the EMA once with mean=None -> zeros, in training mode (Q4)
        self.ema_step = 0
        self.ema_mean = torch.zeros(sum(p.numel() for p in self.params.values()))
        self._ema_update()
        self.weight_decay, self.momentum, self.trust_coef, self.lars_eps = weight_decay, momentum, trust_coef, lars_eps
        self.momentum_buf = OrderedDict((k, None) for k in self.names)

    # ---- flat views -------------------------------------------------------------------------
    def flat_params(self):
        return torch.cat([p.detach().reshape(-1) for p in self.params.values()])   # parameters_to_vector

    def target_params(self):
        out, off = OrderedDict(), 0
        for k, p in self.params.items():
            out[k] = self.ema_mean[off:off + p.numel()].view_as(p)
            off += p.numel()
        return out

    def _ema_update(self):
        """main.py:159-162."""
        decay = cos_ema_decay(self.ema_step, self.total_steps, self.base_decay)
        self.ema_mean = (1 - decay) * self.flat_params() + decay * self.ema_mean
        self.ema_step += 1

    # ---- forward (main.py:242-276) ----------------------------------------------------------
    def forward(self, aug1, aug2, training=True):
        P = self.params
        q = self.q
        o1 = prediction(self.arch, P, self.bn, aug1, training, q)
        o2 = prediction(self.arch, P, self.bn, aug2, training, q)
        T = self.target_params()
        with torch.no_grad():   # bit-identical to the reference's graph-building target passes (SURVEY.md §3.3 probe)
            t1 = prediction(self.arch, T, self.bn, aug1, training, q)
            t2 = prediction(self.arch, T, self.bn, aug2, training, q)
        rep = torch.cat([o1[0], o2[0]], 0) if training else o1[0]
        linear_preds = F.linear(q(rep.clone().detach()), q(P["linear_classifier.weight"]), P["linear_classifier.bias"])
        if training:
            self._ema_update()
        return {
            "linear_preds": linear_preds,
            "online_representation1": o1[0], "online_projection1": o1[1], "online_prediction1": o1[2],
            "online_representation2": o2[0], "online_projection2": o2[1], "online_prediction2": o2[2],
            "target_representation1": t1[0], "target_projection1": t1[1], "target_prediction1": t1[2],
            "target_representation2": t2[0], "target_projection2": t2[1], "target_prediction2": t2[2],
        }

    # ---- one optimisation step (main.py:589-624), `world` emulated ranks --------------------
    def train_step(self, aug1, aug2, labels, lr, world=1, sync_bn=False):
        """aug*: [world*b, 3, R, R]; rank r owns rows r*b:(r+1)*b.  Returns a dict of detached scalars/tensors.
        world == 1 is exactly the single-process reference.  world > 1 needs sync_bn=True (global BN statistics,
        which is what SyncBatchNorm computes with equal per-rank counts) to be expressible as one graph."""
        if world > 1 and not sync_bn:
            raise NotImplementedError("rank-local BN under DDP: run one OracleBYOL per rank and average grads")
        for p in self.params.values():
            p.grad = None
        out = self.forward(aug1, aug2, training=True)
        b = aug1.shape[0] // world
        byol_losses, ce_losses = [], []
        lab2 = None
        for r in range(world):
            sl = slice(r * b, (r + 1) * b)
            byol_losses.append(loss_function(out["online_prediction1"][sl], out["online_prediction2"][sl],
                                             out["target_projection1"][sl], out["target_projection2"][sl]))
            lp = torch.cat([out["linear_preds"][sl], out["linear_preds"][world * b + r * b: world * b + (r + 1) * b]], 0)
            lab2 = torch.cat([labels[sl], labels[sl]], 0)
            ce_losses.append(F.cross_entropy(lp, lab2))
        byol_loss = torch.stack(byol_losses).mean()     # DDP: gradients are averaged over ranks
        ce_loss = torch.stack(ce_losses).mean()
        loss = byol_loss + ce_loss
        loss.backward()
        grads = OrderedDict((k, p.grad.detach().clone()) for k, p in self.params.items())
        self.lars_sgd_step(lr)
        res = {k: v.detach() for k, v in out.items()}
        res.update(loss=loss.detach(), byol_loss=byol_loss.detach(), ce_loss=ce_loss.detach(), grads=grads)
        return res

    # ---- LARS (optimizers/lars.py:84-127) around SGD(momentum) (main.py:316,336) -------------
    def lars_sgd_step(self, lr):
        with torch.no_grad():
            for k, p in self.params.items():
                if p.grad is None:
                    continue
                g = p.grad
                ignore = is_ignored(k, p)
                wd = 0.0 if ignore else self.weight_decay
                if wd > 0:
                    g = g.add(p, alpha=wd)
                if not ignore:
                    param_norm, grad_norm = p.norm(), g.norm()
                    adaptive_lr = 1.0
                    if param_norm > 0 and grad_norm > 0:
                        adaptive_lr = self.trust_coef * param_norm / (grad_norm + self.lars_eps)
                    g = g.mul(adaptive_lr)
                buf = self.momentum_buf[k]
                if buf is None:
                    buf = torch.clone(g).detach()
                else:
                    buf.mul_(self.momentum).add_(g, alpha=1)
                self.momentum_buf[k] = buf
                p.add_(buf, alpha=-lr)
