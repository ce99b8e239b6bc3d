"""Debug aid (GPU): per-block relative error of the CUDA train-mode forward vs the CPU oracle."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import byol_oracle as O
from byol_b200.model import BYOL

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
r = int(sys.argv[3]) if len(sys.argv) > 3 else 64
train = (sys.argv[4] != "eval") if len(sys.argv) > 4 else True
seed = 11
torch.manual_seed(seed)
rep = 512 if arch in ("resnet18", "resnet34") else 2048
model = BYOL(rep, 256, 1000, 10, arch=arch).cuda()
model.train(train)
params, buffers = O.init_reference_state(arch, seed)
storage = sys.argv[5] if len(sys.argv) > 5 else "fp32"
oracle = O.OracleBYOL(arch, params, buffers, 10, storage=storage)
g = torch.Generator().manual_seed(5)
a1, a2 = torch.rand(b, 3, r, r, generator=g), torch.rand(b, 3, r, r, generator=g)
trace = []
with torch.no_grad():
    rep_ref = O.encoder_forward(arch, oracle.params, oracle.bn, a1, train, trace=trace, q=oracle.q)
eng = model._ensure_ready(b)
eng.prep_weights(eng.theta, eng.w_online, want_dgrad=False)
saved = {}
with torch.no_grad():
    outs, _ = eng.forward_lanes([a1.cuda()], [(eng.theta, eng.w_online, saved)], train)
torch.cuda.synchronize()
def rel(a, b):
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())
print(arch, "b", b, "r", r, "train", train, "oracle storage", storage)
tr = dict(trace[:3])
y0 = saved["y0"].float().cpu().permute(0, 3, 1, 2)
print("stem conv y0           max-rel %.3e  l2-rel %.3e" % rel(y0, tr["stem_conv"]))
c0 = saved["c0"].cpu()
y0r = tr["stem_conv"]
print("stem mean err %.3e invstd relerr %.3e" % (float((c0[2] - y0r.mean((0, 2, 3))).abs().max()),
      float((c0[3] * torch.sqrt(y0r.var((0, 2, 3), unbiased=False) + 1e-5) - 1).abs().max())))
x_in = saved["blocks"][0]["x"].float().cpu().permute(0, 3, 1, 2)
print("pool out               max-rel %.3e  l2-rel %.3e" % rel(x_in, tr["pool"]))
b0 = saved["blocks"][0]
for (name, ref), blk in zip(trace[3:], saved["blocks"]):
    got = blk["out"].float().cpu().permute(0, 3, 1, 2)
    print("%-22s max-rel %.3e  l2-rel %.3e" % (name, *rel(got, ref)))
print("representation        max-rel %.3e  l2-rel %.3e" % rel(outs[0][0].cpu(), rep_ref))
