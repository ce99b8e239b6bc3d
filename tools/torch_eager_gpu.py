"""Context measurement (NOT the product, NOT bench.py): the reference's step semantics executed with stock PyTorch
modules on the GPU (cuDNN / cuBLAS / ATen, channels_last, optional autocast bf16), i.e. what running
/root/reference/main.py on torch 2.11 would do on this box.  Restates main.py:214-276 (param-swap target passes WITH
autograd graph, Q6), objective.py:6-25, optimizers/lars.py:84-127 (per-tensor Python loop with host syncs).

    python tools/torch_eager_gpu.py --batch 256 --steps 5 --autocast 1
"""
import argparse
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import byol_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--autocast", type=int, default=1)
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--target-no-grad", type=int, default=0, help="1 = evaluate target passes under no_grad (not what the reference does)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True        # main.py:705
    torch.manual_seed(0)
    m = O.build_reference_modules(a.arch, O.representation_size(a.arch)).to(dev)
    m.train()
    params = list(m.parameters())
    mean = 0.004 * nn.utils.parameters_to_vector(params).detach()
    ema_step = 1
    groups = [{"params": [p for n, p in m.named_parameters() if p.dim() == 1 or n.endswith(".bias")], "weight_decay": 0.0, "ignore": True},
              {"params": [p for n, p in m.named_parameters() if not (p.dim() == 1 or n.endswith(".bias"))], "weight_decay": 1e-6, "ignore": False}]
    opt = torch.optim.SGD(groups, lr=0.2 * a.batch / 256, momentum=0.9)

    def prediction(x):
        r = m.base_network(x).view(-1, 2048 if O.representation_size(a.arch) == 2048 else 512)
        pj = m.head(r)
        return r, pj, m.predictor(pj)

    def target_prediction(x):
        orig = nn.utils.parameters_to_vector(params)
        nn.utils.vector_to_parameters(mean, params)
        if a.target_no_grad:
            with torch.no_grad():
                out = prediction(x)
        else:
            out = prediction(x)
        nn.utils.vector_to_parameters(orig, params)
        return out

    def reg(x, y):
        return -2 * torch.sum(x * y, dim=-1) / (x.norm() * y.norm())

    def lars():
        with torch.no_grad():
            for g in opt.param_groups:
                for p in g["params"]:
                    if p.grad is None:
                        continue
                    if g["weight_decay"] > 0:
                        p.grad = p.grad.add(p, alpha=g["weight_decay"])
                    if not g["ignore"]:
                        pn, gn = p.norm(), p.grad.norm()
                        alr = 1.0
                        if pn > 0 and gn > 0:        # host sync x2 per tensor, as in lars.py:107
                            alr = 0.001 * pn / gn
                        p.grad = p.grad.mul(alr)
        wds = [g["weight_decay"] for g in opt.param_groups]
        for g in opt.param_groups:
            g["weight_decay"] = 0
        opt.step()
        for g, w in zip(opt.param_groups, wds):
            g["weight_decay"] = w

    g = torch.Generator(device=dev).manual_seed(1234)
    x1 = torch.rand(a.batch, 3, 224, 224, generator=g, device=dev)
    x2 = torch.rand(a.batch, 3, 224, 224, generator=g, device=dev)
    lab = torch.randint(0, 1000, (a.batch,), generator=g, device=dev)

    def step():
        nonlocal mean, ema_step
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bool(a.autocast)):
            o1, o2 = prediction(x1), prediction(x2)
            t1, t2 = target_prediction(x1), target_prediction(x2)
            lp = m.linear_classifier(torch.cat([o1[0], o2[0]], 0).clone().detach())
            d = 1 - (1 - 0.996) * (np.cos(np.pi * ema_step / 1000) + 1) / 2.0
            mean = (1 - d) * nn.utils.parameters_to_vector(params).detach() + d * mean
            ema_step += 1
            loss = torch.mean(reg(o1[2].float(), t2[1].detach().float()) + reg(o2[2].float(), t1[1].detach().float()))
            loss = loss + F.cross_entropy(lp.float(), torch.cat([lab, lab], 0))
        opt.zero_grad()
        loss.backward()
        lars()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print({"torch_eager_images_per_sec": a.batch / ms * 1000, "ms_per_step": ms, "batch": a.batch, "autocast_bf16": bool(a.autocast),
           "target_no_grad": bool(a.target_no_grad), "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})


if __name__ == "__main__":
    main()
