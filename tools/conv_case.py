"""Run one conv shape repeatedly (for ncu captures):  python tools/conv_case.py cin cout k stride hin batch kind"""
import sys
import torch
sys.path.insert(0, ".")
from byol_b200 import ops

cin, cout, k, s, hin, b = [int(x) for x in sys.argv[1:7]]
kind = sys.argv[7] if len(sys.argv) > 7 else "fprop"
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
p = k // 2
dev = torch.device("cuda")
cpad = (cin + 7) // 8 * 8
ho = ops.conv_out_size(hin, k, s, p)
x = torch.randn(b, hin, hin, cpad, device=dev).to(torch.bfloat16)
dy = torch.randn(b, ho, ho, cout, device=dev).to(torch.bfloat16)
w = torch.randn(cout, cin, k, k, device=dev) * 0.05
wf, wd = ops.prep_weight(w, cpad=cpad, want_dgrad=cin % 8 == 0)
dw = torch.zeros(cout, cin, k, k, device=dev)
stats = torch.zeros(2 * cout, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(reps):
    flush.fill_(i)
    if kind == "fprop":
        ops.conv_fprop(x, wf, k, k, s, p, stats=stats)
    elif kind == "fprop_nostats":
        ops.conv_fprop(x, wf, k, k, s, p)
    elif kind == "dgrad":
        ops.conv_dgrad(dy, wd, hin, hin, k, k, s, p)
    else:
        ops.conv_wgrad(x, dy, dw, k, k, s, p)
torch.cuda.synchronize()
print("done", kind)
