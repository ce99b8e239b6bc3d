"""GPU parity, teacher-forced per block: each residual block (and the stem) of the CUDA engine is run forward and
backward from ITS OWN saved input, and compared with torch autograd on the oracle's restatement of that block fed
the same bf16 input, the same (bf16-rounded) weights and the same upstream gradient.  One block at a time keeps
BatchNorm conditioning out of the comparison, so tolerances are bf16-level: l2 errors < 5e-3 forward and
< 4e-2 backward (activation gradients are stored in bf16, the oracle keeps them in fp32), parameter-gradient cosine > 0.995; max-norm errors are looser because a bf16 rounding flip next
to a ReLU threshold changes single elements by a full unit.
"""
import copy

import pytest
import torch

from oracle import byol_oracle as O
from byol_b200 import engine as E

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20)), float((a - b).norm() / (b.norm() + 1e-20))


def _nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("arch,rep", [("resnet:bottleneck:2,1,1,1", 2048), ("resnet:basic:2,1,2,1", 512)])
def test_blocks_teacher_forced(cuda, arch, rep):
    from byol_b200.model import BYOL
    seed, b, r = 31, 16, 64
    torch.manual_seed(seed)
    model = BYOL(rep, 256, 1000, 10, arch=arch).cuda().train()
    params, buffers = O.init_reference_state(arch, seed)
    kind, _ = O.arch_spec(arch)
    g = torch.Generator().manual_seed(3)
    a1 = torch.rand(b, 3, r, r, generator=g)
    eng = model._ensure_ready(b)
    eng.prep_weights(eng.theta, eng.w_online, want_dgrad=True)
    saved = {}
    with torch.no_grad():
        eng.forward_lanes([a1.cuda()], [(eng.theta, eng.w_online, saved)], True)
    names = list(params.keys())
    offs, off = {}, 0
    for k in names:
        offs[k] = (off, params[k].numel())
        off += params[k].numel()
    q = O.bf16_storage
    worst = {"fwd": 0.0, "gin": 0.0, "cos": 1.0}

    def check_param_grads(P, label):
        for k, p in P.items():
            if p.grad is None:
                continue
            o, n = offs[k]
            got = eng.grad[o:o + n].cpu().double()
            ref = p.grad.reshape(-1).double()
            cos = float((got @ ref) / (got.norm() * ref.norm() + 1e-30))
            ratio = float(got.norm() / (ref.norm() + 1e-30))
            print("    %-44s cos %.5f  norm ratio %.4f  |g| %.3e" % (k, cos, ratio, float(ref.norm())))
            worst["cos"] = min(worst["cos"], cos)
            assert cos > 0.995 and 0.97 < ratio < 1.03, (label, k, cos, ratio)

    # ---- residual blocks ------------------------------------------------------------------
    for bi, (prefix, stride) in enumerate(O.block_list(arch)):
        S = saved["blocks"][bi]
        blk = eng.blocks[bi]
        gg = torch.Generator().manual_seed(100 + bi)
        g_out = torch.randn(S["out"].shape, generator=gg).to(torch.bfloat16)
        eng.grad.zero_()
        eng._bpool = E._Pool(2 * eng.bn_channels, eng.device, zero=True)
        g_in = eng._block_bwd(blk, [S], [g_out.cuda()])[0]
        torch.cuda.synchronize()
        x = _nchw(S["x"]).requires_grad_(True)
        P = {k: v.clone().requires_grad_(True) for k, v in params.items() if k.startswith(prefix + ".")}
        bn = O._BN(copy.deepcopy(buffers))
        out = O.block_forward(kind, P, bn, x, prefix, stride, True, q)
        out.backward(_nchw(g_out))
        ef, eg = _rel(_nchw(S["out"]), out), _rel(_nchw(g_in), x.grad)
        print("block %s fwd max/l2 %.2e/%.2e  g_in max/l2 %.2e/%.2e" % (prefix, ef[0], ef[1], eg[0], eg[1]))
        worst["fwd"], worst["gin"] = max(worst["fwd"], ef[1]), max(worst["gin"], eg[1])
        assert ef[1] < 5e-3 and ef[0] < 2e-2, prefix
        assert eg[1] < 5e-2 and eg[0] < 0.5, prefix   # gradients are stored in bf16 between kernels; the oracle keeps fp32
        check_param_grads(P, prefix)

    # ---- stem: conv7x7/2 -> BN -> ReLU -> maxpool ---------------------------------------------
    pool_out = saved["blocks"][0]["x"]
    gg = torch.Generator().manual_seed(99)
    g_pool = torch.randn(pool_out.shape, generator=gg).to(torch.bfloat16)
    eng.grad.zero_()
    eng._bpool = E._Pool(2 * eng.bn_channels, eng.device, zero=True)
    from byol_b200 import ops
    n, h, w, c = saved["a0_shape"]
    g0 = ops.maxpool_bwd(g_pool.cuda(), saved["pool_idx"], h, w, eng.pool_k, eng.pool_s, eng.pool_p)
    dy0, _ = eng._bn_bwd(eng.stem, [g0], [saved["y0"]], [saved["c0"]], 1)
    eng._wgrad(eng.stem, [saved["x8"]], dy0)
    torch.cuda.synchronize()
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()
         if k in ("base_network.0.weight", "base_network.1.weight", "base_network.1.bias")}
    bn = O._BN(copy.deepcopy(buffers))
    out = O.stem_forward(P, bn, a1, True, q=q)
    out.backward(_nchw(g_pool))
    ef = _rel(_nchw(pool_out), out)
    print("stem fwd max/l2 %.2e/%.2e" % ef)
    assert ef[1] < 5e-3
    check_param_grads(P, "stem")
    print("worst: fwd l2 %.2e  g_in l2 %.2e  param-grad cosine %.5f" % (worst["fwd"], worst["gin"], worst["cos"]))
