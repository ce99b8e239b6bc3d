"""Time conv shapes with CUDA events (L2 flushed between launches):
   python tools/time_cases.py "cin cout k s hin batch kind" ...   kind: fprop | fprop_nostats | dgrad | wgrad"""
import sys
import torch
sys.path.insert(0, ".")
from byol_b200 import ops

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for spec in sys.argv[1:]:
    f = spec.split()
    cin, cout, k, s, hin, b = [int(x) for x in f[:6]]
    kind = f[6]
    p = k // 2
    cpad = (cin + 7) // 8 * 8
    ho = ops.conv_out_size(hin, k, s, p)
    x = torch.randn(b, hin, hin, cpad, device=dev).to(torch.bfloat16)
    dy = torch.randn(b, ho, ho, cout, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    wf, wd = ops.prep_weight(w, cpad=cpad, want_dgrad=cin % 8 == 0)
    dw = torch.zeros(cout, cin, k, k, device=dev)
    stats = torch.zeros(2 * cout, device=dev)
    ts = []
    for i in range(12):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if kind == "fprop":
            ops.conv_fprop(x, wf, k, k, s, p, stats=stats)
        elif kind == "fprop_nostats":
            ops.conv_fprop(x, wf, k, k, s, p)
        elif kind == "dgrad":
            ops.conv_dgrad(dy, wd, hin, hin, k, k, s, p)
        else:
            ops.conv_wgrad(x, dy, dw, k, k, s, p)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    print(f"{spec:40s} median {ts[len(ts)//2]*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us", flush=True)
