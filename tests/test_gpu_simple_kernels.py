"""GPU parity for the HBM-bound kernels: BN fwd/bwd, pooling, layout, loss, EMA (bit-exact), LARS+SGD."""
import numpy as np
import pytest
import torch

from oracle import ops_ref as R
from tests.util import assert_close

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.mark.parametrize("m,c", [(300, 64), (128, 256), (70, 4096), (1000, 8)])
def test_bn_forward(cuda, m, c):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = R.bf16_round(torch.randn(m, c, generator=g) * 2 + 0.5)
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g)
    rm, rv = torch.zeros(c), torch.ones(c)
    xd = x.to(cuda, BF)
    stats = torch.zeros(2 * c, device=cuda)
    ops.bn_stats(xd, stats)
    coeffs = torch.empty(4, c, device=cuda)
    rmd, rvd = rm.to(cuda), rv.to(cuda)
    ops.bn_finalize(stats, m, gamma.to(cuda), beta.to(cuda), rmd, rvd, 0.1, 1e-5, coeffs)
    y = ops.bn_apply(xd, coeffs[0], coeffs[1], relu=True)
    torch.cuda.synchronize()
    yref, mean, invstd, var = R.bn_train_ref(x, gamma, beta)
    assert_close("bn_mean", coeffs[2], mean, atol=1e-5, rtol=1e-5)
    assert_close("bn_invstd", coeffs[3], invstd, atol=0, rtol=1e-4)
    assert_close("bn_apply_relu", y, torch.relu(yref), atol=2e-2, rtol=1e-2)
    assert_close("bn_running_mean", rmd, 0.9 * rm + 0.1 * mean, atol=1e-5, rtol=1e-5)
    assert_close("bn_running_var", rvd, 0.9 * rv + 0.1 * var * m / (m - 1), atol=1e-5, rtol=1e-4)


def test_bn_finalize_lanes_matches_sequential(cuda):
    """The multi-lane finalize equals four sequential single-lane calls (coefficients and running statistics)."""
    from byol_b200 import ops
    c, m, L = 256, 500, 4
    g = torch.Generator().manual_seed(11)
    stats = (torch.rand(L * 2 * c, generator=g) * 100).to(cuda)
    stats.view(L, 2, c)[:, 1] += 200.0     # make E[x^2] >= E[x]^2
    gam = [torch.rand(c, generator=g).to(cuda) + 0.5 for _ in range(L)]
    bet = [torch.randn(c, generator=g).to(cuda) for _ in range(L)]
    rm1, rv1 = torch.zeros(c, device=cuda), torch.ones(c, device=cuda)
    rm2, rv2 = rm1.clone(), rv1.clone()
    co1 = torch.empty(L, 4, c, device=cuda)
    co2 = torch.empty(L, 4, c, device=cuda)
    for l in range(L):
        ops.bn_finalize(stats[l * 2 * c:(l + 1) * 2 * c], m, gam[l], bet[l], rm1, rv1, 0.1, 1e-5, co1[l])
    ops.bn_finalize_lanes(stats, m, gam, bet, rm2, rv2, 0.1, 1e-5, co2)
    torch.cuda.synchronize()
    # same arithmetic, separately compiled kernels (FMA contraction may differ): agree to an ulp or two
    assert torch.allclose(co1, co2, rtol=1e-6, atol=1e-7)
    assert torch.allclose(rm1, rm2, rtol=1e-6, atol=1e-7) and torch.allclose(rv1, rv2, rtol=1e-6, atol=1e-7)


def test_bn_apply_residual(cuda):
    from byol_b200 import ops
    m, c = 200, 128
    g = torch.Generator().manual_seed(2)
    x = R.bf16_round(torch.randn(m, c, generator=g))
    r = R.bf16_round(torch.randn(m, c, generator=g))
    sc, sh, rs, rsh = [torch.randn(c, generator=g) for _ in range(4)]
    y1 = ops.bn_apply(x.to(cuda, BF), sc.to(cuda), sh.to(cuda), relu=True, resid=r.to(cuda, BF))
    y2 = ops.bn_apply(x.to(cuda, BF), sc.to(cuda), sh.to(cuda), relu=True, resid=r.to(cuda, BF), rscale=rs.to(cuda),
                      rshift=rsh.to(cuda))
    torch.cuda.synchronize()
    assert_close("bn_apply_resid", y1, torch.relu(x * sc + sh + r), atol=3e-2, rtol=1e-2)
    assert_close("bn_apply_resid_affine", y2, torch.relu(x * sc + sh + r * rs + rsh), atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("mask_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("m,c", [(300, 64), (96, 4096)])
def test_bn_backward(cuda, mask_mode, m, c):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = R.bf16_round(torch.randn(m, c, generator=g) + 0.3)
    gy = R.bf16_round(torch.randn(m, c, generator=g))
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g) * 0.1
    yref, mean, invstd, var = R.bn_train_ref(x, gamma, beta)
    scale = gamma * invstd
    shift = beta - mean * scale
    coeffs = torch.stack([scale, shift, mean, invstd]).to(cuda)
    act = None
    if mask_mode == 0:
        dz = gy
    elif mask_mode == 1:
        dz = gy * ((x * scale + shift) > 0)
    else:
        act = R.bf16_round(torch.relu(yref + torch.randn(m, c, generator=g)))
        dz = gy * (act > 0)
    dxref, dgamma, dbeta = R.bn_bwd_ref(dz, x, mean, invstd, gamma)
    s12 = torch.zeros(2 * c, device=cuda)
    actd = act.to(cuda, BF) if act is not None else None
    if mask_mode == 3:
        # mask bits as bn_apply writes them: out = relu(1*act + 0) = act, bit = out > 0
        bits = torch.empty(m * c // 8, dtype=torch.uint8, device=cuda)
        out = ops.bn_apply(actd, torch.ones(c, device=cuda), torch.zeros(c, device=cuda), relu=True, mask_out=bits)
        torch.cuda.synchronize()
        expect = (act.view(-1, 8) > 0).to(torch.int32) * (2 ** torch.arange(8, dtype=torch.int32))
        assert torch.equal(bits.cpu().to(torch.int32), expect.sum(1)) and torch.equal(out.cpu().float(), act)
        actd = bits
    ops.bn_bwd_reduce(gy.to(cuda, BF), x.to(cuda, BF), coeffs, s12, mask_mode, act=actd)
    dz_out = torch.empty(m, c, device=cuda, dtype=BF)
    dgam, dbet = torch.ones(c, device=cuda), torch.ones(c, device=cuda)
    dy = ops.bn_bwd_apply(gy.to(cuda, BF), x.to(cuda, BF), coeffs, gamma.to(cuda), s12, m, mask_mode, act=actd,
                          dz_out=dz_out, dgamma=dgam, dbeta=dbet)
    torch.cuda.synchronize()
    assert_close("bn_bwd_dgamma_acc", dgam, 1 + dgamma, atol=1e-3 * float(dgamma.abs().max()) + 1e-4, rtol=1e-4)
    assert_close("bn_bwd_dbeta_acc", dbet, 1 + dbeta, atol=1e-3 * float(dbeta.abs().max()) + 1e-4, rtol=1e-4)
    assert_close("bn_bwd_dbeta", s12[:c], dbeta, atol=1e-3 * float(dbeta.abs().max()) + 1e-4, rtol=1e-4)
    assert_close("bn_bwd_dgamma", s12[c:], dgamma, atol=1e-3 * float(dgamma.abs().max()) + 1e-4, rtol=1e-4)
    assert_close("bn_bwd_dx", dy, dxref, atol=1e-2 * float(dxref.abs().max()), rtol=1e-2)
    assert_close("bn_bwd_dz", dz_out, dz, atol=0, rtol=0)


def test_col_sum(cuda):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(333, 1000, generator=g)
    out = torch.zeros(1000, device=cuda)
    ops.col_sum(x.to(cuda), out)
    outb = torch.zeros(1000, device=cuda)
    ops.col_sum(R.bf16_round(x).to(cuda, BF), outb)
    torch.cuda.synchronize()
    assert_close("col_sum_f32", out, x.sum(0), atol=1e-3, rtol=1e-4)
    assert_close("col_sum_bf16", outb, R.bf16_round(x).sum(0), atol=1e-3, rtol=1e-4)


def test_layout_and_weights(cuda):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.rand(3, 3, 10, 12, generator=g)
    y = ops.nchw_to_nhwc8(x.to(cuda))
    torch.cuda.synchronize()
    ref = torch.zeros(3, 10, 12, 8)
    ref[..., :3] = x.permute(0, 2, 3, 1)
    assert_close("nchw_to_nhwc8", y, R.bf16_round(ref), atol=0, rtol=0)
    w = torch.randn(16, 5, 3, 3, generator=g)
    wf, wd = ops.prep_weight(w.to(cuda), cpad=8)
    torch.cuda.synchronize()
    reff = torch.zeros(16, 3, 3, 8)
    reff[..., :5] = w.permute(0, 2, 3, 1)
    assert_close("prep_weight_fprop", wf.view(16, 3, 3, 8), R.bf16_round(reff), atol=0, rtol=0)
    refd = w.permute(1, 2, 3, 0).contiguous()  # [cin, kh, kw, cout]
    assert_close("prep_weight_dgrad", wd.view(5, 3, 3, 16), R.bf16_round(refd), atol=0, rtol=0)


def test_pooling(cuda):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = R.bf16_round(torch.relu(torch.randn(2, 12, 12, 16, generator=g)))  # ties at zero like post-ReLU maps
    y, idx = ops.maxpool_fwd(x.to(cuda, BF))
    torch.cuda.synchronize()
    yref, _ = R.maxpool_ref(x)
    assert_close("maxpool_fwd", y, yref, atol=0, rtol=0)
    dy = R.bf16_round(torch.randn(2, 6, 6, 16, generator=g))
    dx = ops.maxpool_bwd(dy.to(cuda, BF), idx, 12, 12)
    torch.cuda.synchronize()
    assert_close("maxpool_bwd", dx, R.maxpool_bwd_ref(x, dy), atol=2e-2, rtol=1e-2)
    a = R.bf16_round(torch.randn(5, 7, 7, 64, generator=g))
    yf, yb = ops.avgpool_fwd(a.to(cuda, BF))
    torch.cuda.synchronize()
    assert_close("avgpool_fwd", yf, a.mean((1, 2)), atol=1e-5, rtol=1e-5)
    assert_close("avgpool_fwd_bf16", yb, a.mean((1, 2)), atol=1e-2, rtol=1e-2)
    gb = R.bf16_round(torch.randn(5, 64, generator=g))
    gf = torch.randn(5, 64, generator=g)
    dxa = ops.avgpool_bwd(gb.to(cuda, BF), gf.to(cuda), 5, 7, 7, 64)
    torch.cuda.synchronize()
    assert_close("avgpool_bwd", dxa, ((gb + gf) / 49)[:, None, None, :].expand(5, 7, 7, 64), atol=1e-3, rtol=1e-2)


def test_fused_stem_pool_is_bit_identical(cuda):
    """bn_relu_maxpool_fwd == bn_apply(relu) -> maxpool_fwd, values and argmax indices."""
    from byol_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = R.bf16_round(torch.randn(3, 18, 22, 64, generator=g)).to(cuda, BF)
    sc, sh = (torch.rand(64, generator=g) + 0.5).to(cuda), torch.randn(64, generator=g).to(cuda)
    a = ops.bn_apply(x.view(-1, 64), sc, sh, relu=True).view(3, 18, 22, 64)
    y1, i1 = ops.maxpool_fwd(a)
    y2, i2 = ops.bn_relu_maxpool_fwd(x, sc, sh)
    y3, i3 = ops.bn_relu_maxpool_fwd(x, sc, sh, want_idx=False)     # target lanes: packed-max kernel without argmax
    torch.cuda.synchronize()
    assert torch.equal(y1, y2) and torch.equal(i1, i2)
    assert i3 is None and torch.equal(y1, y3)


def test_subsample2(cuda):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(14)
    x = torch.randn(3, 10, 14, 64, generator=g).to(cuda, BF)
    y = ops.subsample2(x)
    torch.cuda.synchronize()
    assert torch.equal(y, x[:, ::2, ::2, :].contiguous())


def _loss_ref(q1, q2, z1, z2):
    # /root/reference/objective.py:6-25 restated (Frobenius norms of the whole matrices, no per-row normalisation)
    def reg(x, y):
        return -2 * torch.sum(x * y, dim=-1) / (x.norm() * y.norm())
    return torch.mean(reg(q1, z2.detach()) + reg(q2, z1.detach()))


@pytest.mark.parametrize("rows", [8, 256, 512])
def test_loss(cuda, rows):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(7)
    q1, q2, z1, z2 = [torch.randn(rows, 256, generator=g) for _ in range(4)]
    q1r, q2r = q1.clone().requires_grad_(True), q2.clone().requires_grad_(True)
    lref = _loss_ref(q1r, q2r, z1, z2)
    (lref * 0.7).backward()
    d = [t.to(cuda) for t in (q1, q2, z1, z2)]
    ws = torch.empty(6, dtype=torch.float64, device=cuda)
    loss = torch.empty(1, device=cuda)
    saved = torch.empty(6, device=cuda)
    ops.loss_fwd(*d, ws, loss, saved)
    dq1, dq2 = torch.empty_like(d[0]), torch.empty_like(d[1])
    go = torch.tensor([0.7], device=cuda)
    ops.loss_bwd(*d, saved, go, dq1, dq2)
    torch.cuda.synchronize()
    assert_close("loss", loss, lref.detach().reshape(1), atol=1e-7, rtol=1e-5)
    assert_close("loss_dq1", dq1, q1r.grad, atol=1e-9, rtol=1e-4)
    assert_close("loss_dq2", dq2, q2r.grad, atol=1e-9, rtol=1e-4)


@pytest.mark.parametrize("n", [1001, 4096, 1 << 20])
def test_ema_bit_exact(cuda, n):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(n + 4, generator=g)[:n].contiguous()
    mean = torch.randn(n, generator=g)
    # /root/reference/main.py:159-161: decay in float64 numpy, then (1 - decay) * x + decay * mean on fp32 tensors
    step, total, base = 3, 1000, 0.996
    decay = 1 - (1 - base) * (np.cos(np.pi * step / total) + 1) / 2.0
    ref = (1 - decay) * x + decay * mean
    md = mean.to(cuda)
    ops.ema_update(x.to(cuda), md, np.float32(1 - decay), np.float32(decay))
    torch.cuda.synchronize()
    assert torch.equal(md.cpu(), ref), "EMA not bit-exact: max diff %g" % float((md.cpu() - ref).abs().max())


def test_lars_sgd(cuda):
    from byol_b200 import ops
    g = torch.Generator().manual_seed(9)
    shapes = [(64, 3, 7, 7), (64,), (64,), (128, 64, 3, 3), (128,), (1000, 300), (1000,), (70000,)]
    ignore = [0, 1, 1, 0, 1, 0, 1, 1]
    params = [torch.randn(s, generator=g) * 0.1 for s in shapes]
    params[4].zero_()  # zero-norm tensor: adaptive lr falls back to 1 (irrelevant for ignored, but exercise)
    grads = [torch.randn(s, generator=g) * 0.01 for s in shapes]
    lr, wd, trust, mom = 0.3, 1e-6, 0.001, 0.9
    flat_p = torch.cat([p.reshape(-1) for p in params]).to(cuda)
    flat_g = torch.cat([x.reshape(-1) for x in grads]).to(cuda)
    flat_m = torch.zeros_like(flat_p)
    CH = 4096
    cs, cl, ct, first = [], [], [], []
    off = 0
    for t, p in enumerate(params):
        nel = p.numel()
        first.append(len(cs))
        for s in range(0, nel, CH):
            cs.append(s); cl.append(min(CH, nel - s)); ct.append(t)
        off += nel
    first.append(len(cs))
    offs = [0]
    for p in params:
        offs.append(offs[-1] + p.numel())
    table = {
        "p_ptrs": torch.tensor([flat_p.data_ptr() + 4 * o for o in offs[:-1]], dtype=torch.int64, device=cuda),
        "g_ptrs": torch.tensor([flat_g.data_ptr() + 4 * o for o in offs[:-1]], dtype=torch.int64, device=cuda),
        "m_ptrs": torch.tensor([flat_m.data_ptr() + 4 * o for o in offs[:-1]], dtype=torch.int64, device=cuda),
        "chunk_start": torch.tensor(cs, dtype=torch.int64, device=cuda),
        "chunk_len": torch.tensor(cl, dtype=torch.int32, device=cuda),
        "chunk_tensor": torch.tensor(ct, dtype=torch.int32, device=cuda),
        "wd": torch.tensor([0.0 if i else wd for i in ignore], device=cuda),
        "lr": torch.full((len(shapes),), lr, device=cuda),
        "ignore": torch.tensor(ignore, dtype=torch.int32, device=cuda),
        "tensor_first_chunk": torch.tensor(first, dtype=torch.int32, device=cuda),
        "partial": torch.zeros(2 * len(cs), dtype=torch.float64, device=cuda),
    }
    # reference: /root/reference/optimizers/lars.py:84-127 around torch.optim.SGD(momentum=0.9)
    ref_p = [p.clone() for p in params]
    ref_m = [None] * len(params)
    for step in range(2):
        ops.lars_sgd_step(table, trust, 0.0, mom, first_step=(step == 0))
        for i, p in enumerate(ref_p):
            gr = grads[i].clone()
            if not ignore[i]:
                gr = gr.add(p, alpha=wd)
                pn, gn = p.norm(), gr.norm()
                alr = 1.0
                if pn > 0 and gn > 0:
                    alr = trust * pn / (gn + 0.0)
                gr = gr.mul(alr)
            ref_m[i] = gr.clone() if ref_m[i] is None else ref_m[i].mul(mom).add(gr)
            p.add_(ref_m[i], alpha=-lr)
    torch.cuda.synchronize()
    assert_close("lars_params", flat_p, torch.cat([p.reshape(-1) for p in ref_p]), atol=1e-7, rtol=1e-5)
    assert_close("lars_momentum", flat_m, torch.cat([m.reshape(-1) for m in ref_m]), atol=1e-9, rtol=1e-5)


def test_lars_sgd_is_deterministic_and_handles_unaligned_tensors(cuda):
    """Replicas must stay bit-identical under data parallelism (main.py:440): the per-tensor norms are reduced in a
    fixed order (no atomics), so two executions on the same inputs give bit-identical parameters.  Tensors whose
    storage is not 16-byte aligned take the scalar path and must give the same values as the vector path."""
    from byol_b200.lars import LARS
    g = torch.Generator().manual_seed(19)
    shapes = [(257, 129), (3,), (1000, 50), (77,), (300001,)]
    runs = []
    for offset in (0, 0, 1):          # third run: every tensor starts one float off a 16-byte boundary
        torch.manual_seed(0)
        ps, gs = [], []
        g = torch.Generator().manual_seed(19)
        for sh in shapes:
            n = int(np.prod(sh))
            base = torch.zeros(n + 4, device=cuda)
            gbase = torch.zeros(n + 4, device=cuda)
            p = torch.nn.Parameter(base[offset:offset + n].view(sh))
            with torch.no_grad():
                p.copy_(torch.randn(sh, generator=g) * 0.1)
            p.grad = gbase[offset:offset + n].view(sh)
            p.grad.copy_(torch.randn(sh, generator=g) * 0.01)
            ps.append(p); gs.append(p.grad)
        groups = [{"params": [ps[1], ps[3]], "weight_decay": 0.0, "ignore": True},
                  {"params": [ps[0], ps[2], ps[4]], "weight_decay": 1e-6, "ignore": False}]
        opt = LARS(torch.optim.SGD(groups, lr=0.3, momentum=0.9), eps=0.0)
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        runs.append(torch.cat([p.detach().reshape(-1).cpu() for p in ps]))
    assert torch.equal(runs[0], runs[1]), "LARS step is not run-to-run deterministic"
    assert torch.allclose(runs[0], runs[2], rtol=1e-6, atol=1e-8)


def test_lars_rejects_non_fp32(cuda):
    from byol_b200.lars import LARS
    p = torch.nn.Parameter(torch.randn(8, 8, device=cuda, dtype=torch.float16))
    p.grad = torch.randn_like(p)
    opt = LARS(torch.optim.SGD([{"params": [p], "weight_decay": 0.0, "ignore": False}], lr=0.1, momentum=0.9))
    with pytest.raises(NotImplementedError):
        opt.step()


@pytest.mark.parametrize("rows,classes,tile", [(16, 1000, 1), (1024, 1000, 2), (24, 10, 2), (7, 37, 1)])
def test_cross_entropy_topk(cuda, rows, classes, tile):
    """main.py:596-598: F.cross_entropy + metrics.topk((1, 5)) on the classifier logits, here one kernel (+ its
    backward); `tile` = 2 checks the implicit cat([labels, labels])."""
    from byol_b200.objective import cross_entropy_topk
    from byol_b200 import wiring
    g = torch.Generator().manual_seed(23)
    logits = torch.randn(rows, classes, generator=g) * 3
    labels = torch.randint(0, classes, (rows // tile,), generator=g)
    full = torch.cat([labels] * tile)
    logits[0, full[0]] = logits[0].max() + 1.0          # one certain top-1 hit
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr, full)
    (ref * 0.37).backward()
    _, pred = logits.topk(min(5, classes), 1, True, True)
    correct = pred.t().eq(full.view(1, -1))
    ref1 = correct[:1].reshape(-1).float().sum() * 100.0 / rows
    ref5 = correct[:5].reshape(-1).float().sum() * 100.0 / rows
    ld = logits.to(cuda).requires_grad_(True)
    loss, a1, a5 = cross_entropy_topk(ld, labels.to(cuda))
    (loss * 0.37).backward()
    torch.cuda.synchronize()
    assert_close("ce_loss", loss.detach().reshape(1), ref.detach().reshape(1), atol=1e-6, rtol=1e-5)
    assert_close("ce_dlogits", ld.grad, lr.grad, atol=1e-8, rtol=1e-4)
    assert abs(float(a1) - float(ref1)) < 1e-4 and abs(float(a5) - float(ref5)) < 1e-4
    t1, t5 = wiring.topk(logits.to(cuda), full.to(cuda), topk=(1, 5))
    assert abs(float(t1) - float(ref1)) < 1e-4 and abs(float(t5) - float(ref5)) < 1e-4
    # deterministic: the row reduction has a fixed order
    loss2, _, _ = cross_entropy_topk(ld.detach(), labels.to(cuda))
    assert torch.equal(loss2, loss.detach())


@pytest.mark.parametrize("b,k1,h,o", [(512, 2048, 4096, 256), (24, 256, 4096, 256), (200, 512, 256, 64)])
def test_mlp_fused_forward(cuda, b, k1, h, o):
    """main.py:194-205: Linear -> BatchNorm1d -> ReLU -> Linear as ONE cooperative kernel (train and eval mode),
    against fp32 torch on the bf16-rounded operands (the hidden activation is rounded to bf16 between the GEMMs, as
    in the unfused kernels)."""
    from byol_b200 import ops
    if not ops.mlp_fused_supported(b, k1, h, o):
        pytest.skip("shape not supported by the cooperative kernel on this device")
    g = torch.Generator().manual_seed(b + h)
    x = R.bf16_round(torch.randn(b, k1, generator=g))
    w1 = R.bf16_round(torch.randn(h, k1, generator=g) / k1 ** 0.5)
    w2 = R.bf16_round(torch.randn(o, h, generator=g) / h ** 0.5)
    b1, b2 = torch.randn(h, generator=g) * 0.1, torch.randn(o, generator=g) * 0.1
    gamma, beta = torch.rand(h, generator=g) + 0.5, torch.randn(h, generator=g) * 0.1
    dev = lambda t, dt=None: t.to(cuda, dt) if dt is not None else t.to(cuda)
    hid = x.double() @ w1.double().t() + b1.double()
    mean, var = hid.mean(0), hid.var(0, unbiased=False)
    a = torch.relu((hid - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double())
    out_ref = R.bf16_round(a.float()).double() @ w2.double().t() + b2.double()
    stats = torch.zeros(2 * h, device=cuda)
    rm, rv = torch.zeros(h, device=cuda), torch.ones(h, device=cuda)
    coeffs = torch.empty(4, h, device=cuda)
    bar = torch.zeros(2, dtype=torch.int32, device=cuda)
    for rep in range(2):          # twice: the grid barrier state must be reusable
        stats.zero_()
        out, hs, as_ = ops.mlp_fused_fwd(dev(x, BF), dev(w1, BF), dev(b1), dev(gamma), dev(beta), dev(w2, BF), dev(b2),
                                         stats, rm, rv, 0.1, 1e-5, b, coeffs, bar, True, True)
    torch.cuda.synchronize()
    assert_close("mlp_fused_mean", coeffs[2], mean.float(), atol=2e-4, rtol=1e-4)
    assert_close("mlp_fused_invstd", coeffs[3], (1 / torch.sqrt(var + 1e-5)).float(), atol=0, rtol=2e-4)
    assert_close("mlp_fused_h", hs, hid.float(), atol=2e-2, rtol=1e-2)
    assert_close("mlp_fused_a", as_, a.float(), atol=3e-2, rtol=1e-2)
    assert_close("mlp_fused_out", out, out_ref.float(), atol=3e-3 * float(out_ref.abs().max()), rtol=1e-2)
    unb = var * b / (b - 1)
    rm_ref = 0.1 * mean + 0.9 * (0.1 * mean)                   # two updates from (0, 1)
    rv_ref = 0.9 * (0.9 * 1 + 0.1 * unb) + 0.1 * unb
    assert_close("mlp_fused_running_mean", rm, rm_ref.float(), atol=2e-4, rtol=1e-4)
    assert_close("mlp_fused_running_var", rv, rv_ref.float(), atol=1e-5, rtol=2e-4)
    # eval mode: running statistics, nothing saved
    out_e, hs_e, _ = ops.mlp_fused_fwd(dev(x, BF), dev(w1, BF), dev(b1), dev(gamma), dev(beta), dev(w2, BF), dev(b2),
                                       None, rm, rv, 0.1, 1e-5, b, coeffs, bar, False, False)
    torch.cuda.synchronize()
    ae = torch.relu((hid - rm.cpu().double()) / torch.sqrt(rv.cpu().double() + 1e-5) * gamma.double() + beta.double())
    oe = R.bf16_round(ae.float()).double() @ w2.double().t() + b2.double()
    assert hs_e is None
    assert_close("mlp_fused_out_eval", out_e, oe.float(), atol=3e-3 * float(oe.abs().max()), rtol=1e-2)


def test_maxpool_bwd_generic_and_fast_paths_agree(cuda):
    """The 2x2-block kernel (k3 s2 p1, even sizes) and the generic window-search kernel against autograd."""
    from byol_b200 import ops
    g = torch.Generator().manual_seed(66)
    for h in (12, 11, 56):          # 11 -> generic kernel (H != 2 Ho)
        x = R.bf16_round(torch.relu(torch.randn(3, h, h, 64, generator=g)))
        y, idx = ops.maxpool_fwd(x.to(cuda, BF))
        ho = y.shape[1]
        dy = R.bf16_round(torch.randn(3, ho, ho, 64, generator=g))
        dx = ops.maxpool_bwd(dy.to(cuda, BF), idx, h, h)
        torch.cuda.synchronize()
        assert_close("maxpool_bwd_%d" % h, dx, R.maxpool_bwd_ref(x, dy), atol=2e-2, rtol=1e-2)


def test_prep_weights_multi_matches_single(cuda):
    """One-launch conversion of a whole parameter set (tiled transpose) == the per-tensor kernel, for 1x1, 3x3 and
    non-multiple-of-32 shapes, fprop and dgrad layouts."""
    from byol_b200 import ops
    g = torch.Generator().manual_seed(67)
    shapes = [(64, 64, 1), (128, 64, 3), (72, 40, 3), (1000, 2048, 1), (256, 1024, 1)]
    ws = [torch.randn(co, ci, k, k, generator=g) for co, ci, k in shapes]
    flat = torch.cat([w.reshape(-1) for w in ws]).to(cuda)
    nf = sum(w.numel() for w in ws)
    pool_f = torch.zeros(nf, dtype=BF, device=cuda)
    pool_d = torch.zeros(nf, dtype=BF, device=cuda)
    rows, off = [], 0
    for (co, ci, k), w in zip(shapes, ws):
        rows.append([off, off, off, co, ci, ci, k * k, 0])
        off += w.numel()
    desc = torch.tensor(rows, dtype=torch.int64, device=cuda)
    ops.prep_weights_multi(flat, pool_f, pool_d, desc)
    torch.cuda.synchronize()
    off = 0
    for (co, ci, k), w in zip(shapes, ws):
        wf, wd = ops.prep_weight(w.to(cuda))
        n = w.numel()
        assert torch.equal(pool_f[off:off + n].view_as(wf), wf), (co, ci, k)
        assert torch.equal(pool_d[off:off + n].view_as(wd), wd), (co, ci, k)
        off += n
