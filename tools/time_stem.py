"""Time the stem paths at batch N (default 512), 224x224: python tools/time_stem.py [N]"""
import sys
import torch
sys.path.insert(0, ".")
from byol_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
x = torch.rand(n, 3, 224, 224, device=dev)
w = torch.randn(64, 3, 7, 7, device=dev) / 12
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=8):
    ts = []
    for i in range(reps):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2] * 1e3


x8 = ops.nchw_to_nhwc8(x)
wf = ops.prep_weight_fold(w)
stats = torch.zeros(128, device=dev)
print("nchw_to_nhwc8      %8.1f us" % timed(lambda: ops.nchw_to_nhwc8(x)))
print("igemm stem fprop   %8.1f us" % timed(lambda: ops.conv_fprop(x8, wf, 7, 7, 2, 3, stats=stats)))
if hasattr(ops, "stem_conv_fprop"):
    xs4 = ops.nchw_to_stem4(x)
    ws = ops.prep_weight_stem4(w)
    print("nchw_to_stem4      %8.1f us" % timed(lambda: ops.nchw_to_stem4(x)))
    print("stem4 fprop        %8.1f us" % timed(lambda: ops.stem_conv_fprop(xs4, ws, 224, 224, stats=stats)))
dy = torch.randn(n, 112, 112, 64, device=dev).to(torch.bfloat16)
dw = torch.zeros(64, 3, 7, 7, device=dev)
print("igemm stem wgrad   %8.1f us" % timed(lambda: ops.conv_wgrad(x8, dy, dw, 7, 7, 2, 3)))
if hasattr(ops, "stem_conv_wgrad"):
    print("stem4 wgrad        %8.1f us" % timed(lambda: ops.stem_conv_wgrad(xs4, dy, dw, 224, 224)))
