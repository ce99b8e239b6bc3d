"""GPU: the fp32-accurate forward path (precision="fp32": exact 3-way bf16 splits of every operand, 6 tensor-core
product terms, fp32 conv outputs, fp64 BatchNorm statistics — csrc/split.cu) against

* plain fp32 torch on the CPU for single layers (tight: 1e-5), and
* the UNMODIFIED reference's golden vectors (tests/golden/*.npz, made by make_golden.py from /root/reference's
  main.execute_graph): forward outputs, BYOL loss, CE loss, EMA — north_star's "within 1e-3 relative fp32".
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.test_oracle_golden import _batches, _sample_index, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,tol", [(6, 1e-5), (3, 2e-4)])
@pytest.mark.parametrize("cin,cout,k,stride,hw", [(64, 64, 1, 1, 14), (64, 128, 3, 1, 10), (128, 128, 3, 2, 12),
                                                  (3, 64, 7, 2, 32), (256, 512, 1, 2, 8)])
def test_split_conv_matches_fp32(cuda, T, tol, cin, cout, k, stride, hw):
    """One convolution through the tensor cores with split operands equals F.conv2d in fp32 (NOT bf16-rounded
    inputs): the split is exact and all products are exact in the fp32 accumulator."""
    from byol_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, pad = 3, k // 2
    x = torch.randn(n, cin, hw, hw, generator=g) * 2 + 0.7            # large mean: what BatchNorm inputs look like
    w = torch.randn(cout, cin, k, k, generator=g) * 0.1
    ref = F.conv2d(x.double(), w.double(), None, stride, pad).float().permute(0, 2, 3, 1).contiguous()
    cpad = (cin + 7) // 8 * 8
    xp = ops.nchw_to_planes(x.to(cuda), T, cpad) if cin < 8 else \
        ops.split_planes(x.permute(0, 2, 3, 1).contiguous().view(-1, cin).to(cuda), T)[0].view(n, hw, hw, T * cin)
    wp = torch.empty(cout, k * k * T * cpad, dtype=torch.bfloat16, device=cuda)
    ops.prep_weight_planes(w.to(cuda), T, cpad, wp)
    y = ops.conv_fprop(xp, wp, k, k, stride, pad, out_fp32=True)
    torch.cuda.synchronize()
    err = float((y.cpu() - ref).abs().max() / ref.abs().max())
    print("split conv T=%d %dx%d/%d %d->%d: max rel err %.2e" % (T, k, k, stride, cin, cout, err))
    assert err < tol, err


def test_fp32_statistics_and_apply(cuda):
    """fp64-accumulated BatchNorm statistics + the fused apply / split kernel against torch (double) on the CPU."""
    from byol_b200 import ops
    g = torch.Generator().manual_seed(6)
    m, c, T = 1000, 64, 6
    y = torch.randn(m, c, generator=g) * 0.3 + 5.0                     # |mean| >> std: cancellation-prone
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    resid = torch.randn(m, c, generator=g)
    yd = y.to(cuda)
    stats = torch.zeros(2 * c, dtype=torch.float64, device=cuda)
    ops.stats_f32(yd, stats)
    co = torch.empty(1, 4, c, device=cuda)
    rm, rv = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
    ops.bn_finalize_lanes_f64(stats, m, [gamma.to(cuda)], [beta.to(cuda)], rm, rv, 0.1, 1e-5, co)
    o32, pl, cp, mask = ops.bn_apply_f32(yd, co[0, 0], co[0, 1], True, T, resid=resid.to(cuda), want_out32=True,
                                         want_copy=True, want_mask=True)
    torch.cuda.synchronize()
    yd64 = y.double()
    mean, var = yd64.mean(0), yd64.var(0, unbiased=False)
    ref = torch.relu((yd64 - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double() + resid.double())
    assert float((co[0, 2].cpu().double() - mean).abs().max()) < 1e-6
    assert float((o32.cpu().double() - ref).abs().max() / ref.abs().max()) < 2e-6
    planes = pl.float().cpu().view(m, T, c)
    # planes 0 (=term 0), 2 (= plane 1) and 5 (= plane 2) add up to the fp32 value exactly (to 2^-24)
    recon = planes[:, 0].double() + planes[:, 2].double() + planes[:, 5].double()
    assert float((recon - o32.cpu().double()).abs().max() / ref.abs().max()) < 2e-7
    assert torch.equal(cp.cpu(), o32.cpu().to(torch.bfloat16))
    bits = np.unpackbits(mask.cpu().numpy(), bitorder="little").reshape(m, c)
    assert np.array_equal(bits.astype(bool), (o32.cpu().numpy() > 0))
    assert torch.allclose(rm.cpu(), 0.1 * mean.float(), rtol=1e-5, atol=1e-6)


GOLDEN_TOL = {"rn18_b8_r64": 1e-3, "rn18_b32_r224": 1e-3, "rn50_b8_r64": 1e-3, "rn50_b16_r224": 1e-3}


@pytest.mark.parametrize("name", sorted(GOLDEN_TOL))
def test_fp32_path_matches_reference_golden(cuda, name):
    """north_star: forward / loss / EMA of the reference on identical inputs within 1e-3 relative (fp32)."""
    from byol_b200.model import BYOL
    from byol_b200.objective import loss_function
    from byol_b200 import wiring
    z, arch, rep, b, r, steps, seed, lr, total = load_case(name)
    torch.manual_seed(seed)
    model = BYOL(rep, 256, 1000, total, arch=arch, precision="fp32").cuda().train()
    opt = wiring.LARS(torch.optim.SGD(wiring.add_weight_decay(model, 1e-6), lr=lr, momentum=0.9), eps=0.0)
    idx = _sample_index(int(z["numel"]))
    tol = GOLDEN_TOL[name]
    a1, a2, lab = _batches(seed, steps, b, r)[0]
    out = model(a1.cuda(), a2.cuda())
    byol = loss_function(online_prediction1=out["online_prediction1"], online_prediction2=out["online_prediction2"],
                         target_projection1=out["target_projection1"], target_projection2=out["target_projection2"])
    ce = F.cross_entropy(out["linear_preds"], torch.cat([lab, lab]).cuda())
    torch.cuda.synchronize()
    worst = 0.0
    for key in ("online_prediction1", "online_projection2", "target_projection1", "target_projection2",
                "online_representation1", "target_representation2"):
        ref = torch.from_numpy(z["s0_" + key])
        err = float((out[key].detach().cpu() - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        print("%s %s: max rel err %.2e" % (name, key, err))
        assert err < tol, (key, err)
    ref = torch.from_numpy(z["s0_linear_preds_head"])
    err = float((out["linear_preds"].detach().cpu()[:, :16] - ref).abs().max() / ref.abs().max())
    assert err < tol, ("linear_preds", err)
    print("%s: byol %.6f (ref %.6f) ce %.6f (ref %.6f) worst output err %.2e" %
          (name, byol.item(), float(z["s0_byol_loss"]), ce.item(), float(z["s0_ce_loss"]), worst))
    assert abs(byol.item() - float(z["s0_byol_loss"])) < tol * abs(float(z["s0_byol_loss"]))
    assert abs(ce.item() - float(z["s0_ce_loss"])) < tol * abs(float(z["s0_ce_loss"]))
    # EMA after the step's update (fp32 elementwise, bit-exact formula) and the step counter
    np.testing.assert_allclose(model.target_network.mean.cpu()[idx].numpy(), z["s0_ema_sample"], rtol=1e-5, atol=1e-8)
    assert model.target_network.step == int(z["s0_ema_step"])
    # BN running statistics of the stem (fp64-accumulated batch statistics, 4 updates per step)
    sd = model.state_dict()
    np.testing.assert_allclose(sd["base_network.1.running_mean"].cpu().numpy(), z["s0_bn1_running_mean"], rtol=1e-4,
                               atol=1e-6)
    np.testing.assert_allclose(sd["base_network.1.running_var"].cpu().numpy(), z["s0_bn1_running_var"], rtol=1e-4,
                               atol=1e-6)
    # the (bf16-operand) backward + LARS still runs from the accurate forward; gradient norm within 3 %
    (byol + ce).backward()
    gflat = model._engine.grad.detach().cpu()
    ratio = float(gflat.double().norm()) / float(z["s0_grad_norm"])
    print("%s: gradient norm ratio vs reference %.4f" % (name, ratio))
    assert 0.9 < ratio < 1.1
    opt.step()
    torch.cuda.synchronize()
    assert torch.isfinite(model._engine.theta).all()


def test_fp32_path_graph_replay(cuda):
    """The split path under CUDA-graph replay: the second and third step (captured / replayed) stay on the
    reference's loss curve."""
    from byol_b200.model import BYOL
    from byol_b200 import wiring
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden",
                                           "curve_rn18_b16_r64.npz"))
    arch, rep, b, r, steps, seed, lr, total = z["config"]
    rep, b, r, seed, lr, total = int(rep), int(b), int(r), int(seed), float(lr), int(total)
    torch.manual_seed(seed)
    model = BYOL(rep, 256, 1000, total, arch=str(arch), precision="fp32").cuda().train()
    opt = wiring.LARS(torch.optim.SGD(wiring.add_weight_decay(model, 1e-6), lr=lr, momentum=0.9), eps=0.0)
    data = _batches(seed, 4, b, r)
    got = []
    for s in range(5):
        a1, a2, lab = data[s % 4]
        got.append(float(wiring.train_step(model, opt, a1.cuda(), a2.cuda(), lab.cuda())["loss_mean"]))
    print("fp32-path losses %s reference %s" % (got, list(z["loss"][:5])))
    assert abs(got[0] - z["loss"][0]) < 1e-3 * abs(z["loss"][0])
    np.testing.assert_allclose(got, z["loss"][:5], rtol=1e-2)
    assert any(v != "warm" for v in model._engine.graphs.values())
