"""Time the fused-BatchNorm GEMM epilogues against the unfused kernels they replace (one bottleneck conv3 shape).
    python tools/time_fused.py [planes] [hw] [batch]"""
import sys
import torch
sys.path.insert(0, ".")
from byol_b200 import ops

p = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 56
b = int(sys.argv[3]) if len(sys.argv) > 3 else 512
dev = torch.device("cuda")
M, K, N = b * hw * hw, p, 4 * p
BF = torch.bfloat16
x = torch.randn(M, K, device=dev).to(BF)
w = (torch.randn(N, K, device=dev) * 0.1).to(BF)
resid = torch.randn(M, N, device=dev).to(BF)
g = torch.randn(M, N, device=dev).to(BF)
y = torch.empty(M, N, device=dev, dtype=BF)
out = torch.empty(M, N, device=dev, dtype=BF)
mask = torch.zeros(M * N // 8, dtype=torch.uint8, device=dev)
stats = torch.zeros(2 * N, device=dev)
sc, sh, rs = torch.rand(N, device=dev) + 0.5, torch.randn(N, device=dev), torch.rand(N, device=dev)
co = torch.stack([sc, sh, torch.randn(N, device=dev), torch.rand(N, device=dev) + 0.5])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return 1000 * ts[len(ts) // 2]


x4 = x.view(b, hw, hw, K)
cases = [
    ("conv3 1x1 + stats (old)", lambda: ops.conv_fprop(x4, w, 1, 1, 1, 0, stats=stats, out=y.view(b, hw, hw, N))),
    ("bn_apply resid+relu+mask (old)", lambda: ops.bn_apply(y, sc, sh, True, resid=resid, out=out, mask_out=mask)),
    ("bn_bwd_reduce mask3 (old)", lambda: ops.bn_bwd_reduce(g, y, co, stats, 3, act=mask)),
    ("bn_bwd_apply mask3 (old)", lambda: ops.bn_bwd_apply(g, y, co, sc, stats, M, 3, act=mask, dy=out)),
    ("fused stats-only", lambda: ops.gemm_fused(x, w, stats=stats, no_store=True)),
    ("fused plain store (no epilogue extras)", lambda: ops.gemm_fused(x, w, out=out)),
    ("fused scale+shift store", lambda: ops.gemm_fused(x, w, out=out, colscale=sc, bias=sh)),
    ("fused apply resid+relu", lambda: ops.gemm_fused(x, w, out=out, colscale=sc, bias=sh, resid=resid, relu=True)),
    ("fused apply resid+relu+mask", lambda: ops.gemm_fused(x, w, out=out, colscale=sc, bias=sh, resid=resid, relu=True, mask_out=mask)),
    ("fused bwd_reduce", lambda: ops.gemm_fused(x, w, colscale=sc, bias=sh, resid=g, resid_mask=mask, stats=stats, bwd_reduce=True)),
    ("fused bwd_apply", lambda: ops.gemm_fused(x, w, out=out, colscale=sc, bias=sh, resid=g, resid_mask=mask, resid_colscale=rs)),
]
print("M=%d K=%d N=%d" % (M, K, N))
for name, fn in cases:
    print("%-44s %8.1f us" % (name, timed(fn)))
