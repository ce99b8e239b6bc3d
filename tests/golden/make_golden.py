"""Generate golden vectors by running the UNMODIFIED reference (/root/reference) on CPU in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference's own step loop `main.execute_graph` (main.py:559-662) drives `main.BYOL`, `objective.loss_function`
and `optimizers.lars.LARS` around `torch.optim.SGD(momentum=0.9)`, built exactly as main.build_optimizer does
(main.py:303-344).  Missing submodules (`helpers`, `datasets`, `tree`) come from oracle/ref_shims.  The GPU box has
no /root/reference, so only the small .npz fixtures travel; tests/test_oracle_golden.py pins oracle/byol_oracle.py
against them, and the GPU tests then compare the CUDA path with that oracle.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

CASES = [
    # name, arch, repr, batch, image size, steps, seed, lr
    ("rn18_b8_r64", "resnet18", 512, 8, 64, 3, 7, 0.3),
    ("rn50_b8_r64", "resnet50", 2048, 8, 64, 2, 11, 0.3),
    # BASELINE.json configs[0] geometry (ResNet-18, bs 32, 224x224) and a ResNet-50 case at 224x224: BN statistics
    # over >= 1568 samples per channel instead of 32 (round-2 additions; only the first two are pinned bit-exactly
    # by test_oracle_golden, these two carry forward outputs / losses for the fp32-accuracy tests)
    ("rn18_b32_r224", "resnet18", 512, 32, 224, 2, 21, 0.3),
    ("rn50_b16_r224", "resnet50", 2048, 16, 224, 1, 23, 0.3),
]
# loss-curve fixture: (name, arch, repr, batch, image size, steps, seed, lr); only scalars are stored
CURVES = [("curve_rn18_b16_r64", "resnet18", 512, 16, 64, 20, 31, 0.3)]
TOTAL_STEPS = 10  # CosEMA total_training_steps (small so the cosine schedule visibly moves between steps)
NSAMPLE = 4096


def batches(case):
    name, arch, rep, b, r, steps, seed, lr = case
    g = torch.Generator().manual_seed(seed + 1000)
    out = []
    for _ in range(steps):
        out.append((torch.rand(b, 3, r, r, generator=g), torch.rand(b, 3, r, r, generator=g),
                    torch.randint(0, 1000, (b,), generator=g)))
    return out


def sample_index(numel, seed=12345):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (NSAMPLE,), generator=g)


def run_case(case):
    name, arch, rep, b, r, steps, seed, lr = case
    sys.argv = ["main.py", "--arch=%s" % arch, "--representation-size=%d" % rep, "--num-replicas=1", "--no-cuda",
                "--batch-size=%d" % b, "--image-size-override=%d" % r, "--debug-step"]
    for m in [k for k in sys.modules if k in ("main", "objective") or k.startswith(("optimizers", "helpers", "datasets", "tree"))]:
        del sys.modules[m]
    sys.path[:0] = [REF, os.path.join(ROOT, "oracle", "ref_shims")]
    import main  # noqa: E402  (argparse runs at import, main.py:119)
    main.args.cuda = False
    main.args.distributed_rank = 0
    torch.manual_seed(seed)
    model = main.BYOL(base_network_output_size=rep, projection_output_size=256, classifier_output_size=1000,
                      total_training_steps=TOTAL_STEPS, base_decay=0.996)
    names = [k for k, _ in model.named_parameters()]
    theta0 = torch.nn.utils.parameters_to_vector(model.parameters()).detach().clone()
    idx = sample_index(theta0.numel())
    rec = {"theta0_sum": theta0.double().sum().item(), "theta0_abssum": theta0.double().abs().sum().item(),
           "theta0_sample": theta0[idx].numpy(), "numel": theta0.numel(), "ntensors": len(names),
           "ema0_sample": model.target_network.mean[idx].numpy(), "ema0_step": model.target_network.step}

    # optimizer exactly as main.build_optimizer (main.py:321-340) with a fixed lr instead of the schedule
    import helpers.layers as layers
    from optimizers.lars import LARS
    groups = layers.add_weight_decay(model, 1e-6)
    opt = LARS(torch.optim.SGD(groups, lr=lr, momentum=0.9), eps=0.0)

    captured = {}
    model.register_forward_hook(lambda mod, inp, out: captured.__setitem__("out", out))
    raw_grads = {}
    for k, p in model.named_parameters():
        p.register_hook(lambda g, k=k: raw_grads.__setitem__(k, g.detach().clone()))

    from objective import loss_function
    for step, (a1, a2, lab) in enumerate(batches(case)):
        raw_grads.clear()
        loss_val = main.execute_graph(1, model, [(a1, a2, lab)], None, optimizer=opt, prefix="train")
        out = captured["out"]
        byol = loss_function(online_prediction1=out["online_prediction1"], online_prediction2=out["online_prediction2"],
                             target_projection1=out["target_projection1"], target_projection2=out["target_projection2"])
        ce = torch.nn.functional.cross_entropy(out["linear_preds"], torch.cat([lab, lab], 0))
        pre = "s%d_" % step
        rec[pre + "loss"] = loss_val
        rec[pre + "byol_loss"] = byol.item()
        rec[pre + "ce_loss"] = ce.item()
        for key in ("online_prediction1", "online_projection2", "target_projection1", "target_projection2",
                    "online_representation1", "target_representation2"):
            rec[pre + key] = out[key].detach().numpy()
        rec[pre + "linear_preds_head"] = out["linear_preds"].detach()[:, :16].numpy()
        gflat = torch.cat([raw_grads[k].reshape(-1) if k in raw_grads else torch.zeros(p.numel())
                           for k, p in model.named_parameters()])
        rec[pre + "grad_sample"] = gflat[idx].numpy()
        rec[pre + "grad_norm"] = gflat.double().norm().item()
        theta = torch.nn.utils.parameters_to_vector(model.parameters()).detach()
        rec[pre + "theta_sample"] = theta[idx].numpy()
        rec[pre + "theta_sum"] = theta.double().sum().item()
        rec[pre + "ema_sample"] = model.target_network.mean[idx].numpy()
        rec[pre + "ema_sum"] = model.target_network.mean.double().sum().item()
        rec[pre + "ema_step"] = model.target_network.step
        mom = torch.cat([opt.state[p]["momentum_buffer"].reshape(-1) for p in model.parameters()])
        rec[pre + "momentum_sample"] = mom[idx].numpy()
        sd = model.state_dict()
        rec[pre + "bn1_running_mean"] = sd["base_network.1.running_mean"].numpy().copy()
        rec[pre + "bn1_running_var"] = sd["base_network.1.running_var"].numpy().copy()
        rec[pre + "bn1_num_batches"] = int(sd["base_network.1.num_batches_tracked"])
        rec[pre + "headbn_running_var"] = sd["head.1.running_var"].numpy().copy()[:64]
    rec["param_names"] = np.array(names)
    rec["config"] = np.array([arch, str(rep), str(b), str(r), str(steps), str(seed), str(lr), str(TOTAL_STEPS)])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes; losses", [rec["s%d_loss" % s] for s in range(steps)])
    del sys.path[:2]


def run_curve(case):
    """20 optimisation steps of the unmodified reference on 4 cycling batches: the loss curve the CUDA path must
    follow (main.py:579-631 incl. EMA, LARS, momentum)."""
    name, arch, rep, b, r, steps, seed, lr = case
    sys.argv = ["main.py", "--arch=%s" % arch, "--representation-size=%d" % rep, "--num-replicas=1", "--no-cuda",
                "--batch-size=%d" % b, "--image-size-override=%d" % r, "--debug-step"]
    for m in [k for k in sys.modules if k in ("main", "objective") or k.startswith(("optimizers", "helpers", "datasets", "tree"))]:
        del sys.modules[m]
    sys.path[:0] = [REF, os.path.join(ROOT, "oracle", "ref_shims")]
    import main  # noqa: E402
    main.args.cuda = False
    main.args.distributed_rank = 0
    torch.manual_seed(seed)
    model = main.BYOL(base_network_output_size=rep, projection_output_size=256, classifier_output_size=1000,
                      total_training_steps=steps, base_decay=0.996)
    import helpers.layers as layers
    from optimizers.lars import LARS
    from objective import loss_function
    opt = LARS(torch.optim.SGD(layers.add_weight_decay(model, 1e-6), lr=lr, momentum=0.9), eps=0.0)
    captured = {}
    model.register_forward_hook(lambda mod, inp, out: captured.__setitem__("out", out))
    data = batches((name, arch, rep, b, r, 4, seed, lr))
    losses, byols, ces = [], [], []
    for step in range(steps):
        a1, a2, lab = data[step % len(data)]
        losses.append(main.execute_graph(1, model, [(a1, a2, lab)], None, optimizer=opt, prefix="train"))
        out = captured["out"]
        byols.append(loss_function(online_prediction1=out["online_prediction1"],
                                   online_prediction2=out["online_prediction2"],
                                   target_projection1=out["target_projection1"],
                                   target_projection2=out["target_projection2"]).item())
        ces.append(torch.nn.functional.cross_entropy(out["linear_preds"], torch.cat([lab, lab], 0)).item())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, loss=np.array(losses), byol_loss=np.array(byols), ce_loss=np.array(ces),
                        config=np.array([arch, str(rep), str(b), str(r), str(steps), str(seed), str(lr), str(steps)]))
    print("wrote", path, "losses", ["%.4f" % v for v in losses])
    del sys.path[:2]


def run_lr_schedule():
    """Per-epoch learning rates produced by the reference's own wiring (main.build_lr_schedule, main.py:279-300 +
    optimizers/scheduler.py:4-62) for the default recipe shape (warm-up 10, cosine) and a fixed schedule."""
    out = {}
    for tag, argv in (("cosine_w10_e40", ["--epochs=40", "--warmup=10", "--lr-update-schedule=cosine"]),
                      ("fixed_w3_e12", ["--epochs=12", "--warmup=3", "--lr-update-schedule=fixed"]),
                      ("cosine_w0_e8", ["--epochs=8", "--warmup=0", "--lr-update-schedule=cosine"])):
        sys.argv = ["main.py", "--num-replicas=1", "--no-cuda", "--batch-size=256"] + argv
        for m in [k for k in sys.modules if k in ("main", "objective") or k.startswith(("optimizers", "helpers", "datasets", "tree"))]:
            del sys.modules[m]
        sys.path[:0] = [REF, os.path.join(ROOT, "oracle", "ref_shims")]
        import main  # noqa: E402
        p = torch.nn.Parameter(torch.zeros(4))
        opt = torch.optim.SGD([p], lr=0.2, momentum=0.9)
        sched = main.build_lr_schedule(opt)
        lrs = []
        for _ in range(main.args.epochs):
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            sched.step()
        out[tag] = np.array(lrs, dtype=np.float64)
        out[tag + "_cfg"] = np.array([main.args.epochs, main.args.warmup], dtype=np.int64)
        del sys.path[:2]
    path = os.path.join(HERE, "lr_schedule.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v[:13] for k, v in out.items() if not k.endswith("_cfg")})


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:] 
    sys.argv = sys.argv[:1]
    if not only or "lr" in only:
        run_lr_schedule()
    for c in CURVES:
        if not only or c[0] in only:
            run_curve(c)
    for c in CASES:
        if not only or c[0] in only:
            run_case(c)
