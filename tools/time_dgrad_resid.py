"""Cost of the fused residual add in the 1x1 dgrad epilogue (conv1 of a bottleneck: dy [M, p] -> dx [M, 4p])."""
import sys
import torch
sys.path.insert(0, ".")
from byol_b200 import ops
p = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 56
b = int(sys.argv[3]) if len(sys.argv) > 3 else 512
dev = torch.device("cuda")
BF = torch.bfloat16
dy = torch.randn(b, hw, hw, p, device=dev).to(BF)
wd = (torch.randn(4 * p, p, device=dev) * 0.1).to(BF)          # dgrad layout [Cin=4p, taps*Cout=p]
resid = torch.randn(b, hw, hw, 4 * p, device=dev).to(BF)
mask = torch.randint(0, 255, (b * hw * hw * 4 * p // 8,), device=dev, dtype=torch.uint8)
out = torch.empty(b, hw, hw, 4 * p, device=dev, dtype=BF)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=5):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return 1000 * ts[len(ts) // 2]
print("p=%d hw=%d b=%d" % (p, hw, b))
print("dgrad 1x1 plain            %8.1f us" % timed(lambda: ops.conv_dgrad(dy, wd, hw, hw, 1, 1, 1, 0, out=out)))
print("dgrad 1x1 + resid          %8.1f us" % timed(lambda: ops.conv_dgrad(dy, wd, hw, hw, 1, 1, 1, 0, out=out, resid=resid)))
print("dgrad 1x1 + masked resid   %8.1f us" % timed(lambda: ops.conv_dgrad(dy, wd, hw, hw, 1, 1, 1, 0, out=out, resid=resid, resid_mask=mask)))
