"""GPU end-to-end parity: the byol_b200 training step (BYOL.forward -> loss_function -> backward -> LARS.step)
against (a) golden vectors recorded from the UNMODIFIED reference and (b) the pinned CPU oracle, on the same
seeded inputs and bit-identical initial parameters.

Tolerances (stated, see DESIGN.md "Parity").  The hot path stores activations and tensor-core operands in bf16
(fp32 accumulation, fp32 statistics / loss / optimizer); the reference is fp32 throughout.  Two comparisons:

* TIGHT, against the oracle run with storage="bf16" (the same algorithm with bf16 rounding at the same storage
  points): forward outputs within 2e-2 of the tensor scale, gradient cosine > 0.99.  This is the
  implementation-correctness gate.
* LOOSE, against the golden vectors of the unmodified fp32 reference: bf16 rounding of PRE-BatchNorm conv
  outputs is amplified by mean/std at random initialisation (the error grows ~linearly with depth, identically
  in the CPU bf16-storage oracle and in torch autocast), so only ResNet-18 is asserted (1e-1) and ResNet-50 is
  printed.  Everything that is fp32 in both (loss given its inputs, EMA given theta, LARS given grads) is checked
  at 1e-5 / bit-exact in tests/test_gpu_simple_kernels.py; EMA bookkeeping is checked bit-exactly here.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import byol_oracle as O
from tests.test_oracle_golden import _batches, _sample_index, load_case

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _cos(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("name", ["rn18_b8_r64", "rn50_b8_r64"])
def test_training_steps_match_reference(cuda, name):
    from byol_b200.model import BYOL
    from byol_b200.objective import loss_function
    from byol_b200.lars import LARS
    from oracle.ref_shims.helpers.layers import add_weight_decay

    z, arch, rep, b, r, steps, seed, lr, total = load_case(name)
    torch.manual_seed(seed)
    model = BYOL(rep, 256, 1000, total, arch=arch)                      # same construction order => same init
    theta0 = torch.nn.utils.parameters_to_vector(model.parameters()).detach()
    idx = _sample_index(theta0.numel())
    assert np.array_equal(theta0[idx].numpy(), z["theta0_sample"]), "init differs from the reference"
    assert [k for k, _ in model.named_parameters()] == list(z["param_names"])

    params, buffers = O.init_reference_state(arch, seed)
    oracle = O.OracleBYOL(arch, params, buffers, total, storage="bf16")

    model = model.cuda()
    model.train()
    opt = LARS(torch.optim.SGD(add_weight_decay(model, 1e-6), lr=lr, momentum=0.9), eps=0.0)

    for s, (a1, a2, lab) in enumerate(_batches(seed, steps, b, r)):
        ref = oracle.train_step(a1, a2, lab, lr)
        theta_before = model._engine.theta.clone() if model._engine.device is not None else None
        out = model(a1.cuda(), a2.cuda())
        if s == 0:
            # Q4: construction-time EMA (deferred to first GPU use) then this step's update
            assert model.target_network.step == 2
        byol = loss_function(online_prediction1=out["online_prediction1"], online_prediction2=out["online_prediction2"],
                             target_projection1=out["target_projection1"], target_projection2=out["target_projection2"])
        ce = F.cross_entropy(out["linear_preds"], torch.cat([lab, lab], 0).cuda())
        loss = byol + ce
        opt.zero_grad()
        loss.backward()
        gflat = model._engine.grad.clone()
        opt.step()
        torch.cuda.synchronize()

        pre = "s%d_" % s
        print("step", s, "byol", byol.item(), float(z[pre + "byol_loss"]), ref["byol_loss"].item(),
              "ce", ce.item(), float(z[pre + "ce_loss"]))
        for key in ("online_representation1", "online_projection2", "online_prediction1", "target_projection1",
                    "target_projection2", "target_representation2"):
            e_gold, e_orc = _rel(out[key], torch.from_numpy(z[pre + key])), _rel(out[key], ref[key])
            print("  %-24s rel-err vs fp32 golden %.3e vs bf16-storage oracle %.3e" % (key, e_gold, e_orc))
            assert e_orc < 2e-2, key
            if arch == "resnet18" and s == 0:
                assert e_gold < 1e-1, key
        assert abs(ce.item() - ref["ce_loss"].item()) < 2e-3 * abs(ref["ce_loss"].item())
        assert abs(byol.item() - ref["byol_loss"].item()) < 2e-2 * abs(ref["byol_loss"].item()) + 1e-4
        assert abs(ce.item() - float(z[pre + "ce_loss"])) < 2e-2 * abs(float(z[pre + "ce_loss"]))
        # gradients: direction and norm
        gref = torch.cat([g.reshape(-1) for g in ref["grads"].values()])
        c, nr = _cos(gflat, gref), float(gflat.double().norm().cpu() / gref.double().norm())
        print("  grad cosine %.5f norm ratio %.4f (golden norm %.5g)" % (c, nr, float(z[pre + "grad_norm"])))
        assert c > 0.99 and 0.95 < nr < 1.05
        # per-tensor gradient check for the layers closest to / farthest from the loss
        off = 0
        for k, g in ref["grads"].items():
            n = g.numel()
            if k in ("predictor.3.weight", "head.0.weight", "base_network.0.weight", "linear_classifier.weight",
                     "predictor.1.weight", "base_network.1.bias"):
                ck = _cos(gflat[off:off + n], g)
                print("    grad cos %-28s %.5f" % (k, ck))
                assert ck > 0.98, k
            off += n
        # parameters after the LARS step
        th = model._engine.theta
        upd_c = _cos(th.cpu() - theta0 if s == 0 else th.cpu() - prev_theta, oracle.flat_params() - (theta0 if s == 0 else prev_oracle))
        print("  update cosine %.5f" % upd_c)
        assert upd_c > 0.98
        assert _rel(th[idx.cuda()], oracle.flat_params()[idx]) < 1e-2
        prev_theta, prev_oracle = th.cpu().clone(), oracle.flat_params().clone()
        # EMA: bit-exact bookkeeping w.r.t. our own theta (pre-update), close to the reference
        assert model.target_network.step == int(z[pre + "ema_step"]) == oracle.ema_step
        assert _rel(model.target_network.mean[idx.cuda()], oracle.ema_mean[idx]) < 1e-2
        # BN running statistics: 4 updates per step (Q7)
        sd = model.state_dict()
        assert int(sd["base_network.1.num_batches_tracked"]) == 4 * (s + 1)
        assert _rel(sd["base_network.1.running_mean"], torch.from_numpy(z[pre + "bn1_running_mean"])) < 2e-2
        assert _rel(sd["base_network.1.running_var"], torch.from_numpy(z[pre + "bn1_running_var"])) < 2e-2
        assert _rel(sd["head.1.running_var"], oracle.buffers["head.1.running_var"]) < 3e-2


def test_ema_bookkeeping_bit_exact(cuda):
    """target_network.mean after construction + k forwards equals the reference recurrence applied to OUR theta."""
    from byol_b200.model import BYOL
    torch.manual_seed(3)
    model = BYOL(512, 256, 1000, 10, arch="resnet18").cuda().train()
    g = torch.Generator().manual_seed(5)
    a1, a2 = torch.rand(4, 3, 32, 32, generator=g).cuda(), torch.rand(4, 3, 32, 32, generator=g).cuda()
    with torch.no_grad():
        model(a1, a2)
    theta = torch.nn.utils.parameters_to_vector(model.parameters()).detach().cpu()
    mean = torch.zeros_like(theta)
    for step in range(2):   # construction-time update (step 0) + one forward (step 1); theta unchanged (no optimizer)
        d = O.cos_ema_decay(step, 10, 0.996)
        mean = (1 - d) * theta + d * mean
    assert model.target_network.step == 2
    assert torch.equal(model.target_network.mean.cpu(), mean)
    # flat order / offsets: every parameter is a view of the flat vector at its cumulative offset (Q3)
    off = 0
    for p in model.parameters():
        assert p.data_ptr() == model._engine.theta.data_ptr() + 4 * off
        off += p.numel()


def test_eval_forward(cuda):
    from byol_b200.model import BYOL
    torch.manual_seed(4)
    model = BYOL(512, 256, 1000, 10, arch="resnet18").cuda().eval()
    params, buffers = O.init_reference_state("resnet18", 4)
    oracle = O.OracleBYOL("resnet18", params, buffers, 10)
    g = torch.Generator().manual_seed(6)
    a1, a2 = torch.rand(4, 3, 64, 64, generator=g), torch.rand(4, 3, 64, 64, generator=g)
    with torch.no_grad():
        out = model(a1.cuda(), a2.cuda())
        ref = oracle.forward(a1, a2, training=False)
    assert out["linear_preds"].shape == (4, 1000)            # eval: classifier on view 1 only (main.py:250-251)
    assert model.target_network.step == 1                     # only the construction-time update; eval does not step
    for key in ("online_representation1", "online_prediction2", "target_projection1", "linear_preds"):
        e = _rel(out[key], ref[key])
        print(key, e)
        assert e < 4e-2, key
