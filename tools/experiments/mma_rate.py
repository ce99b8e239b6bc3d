"""python tools/experiments/mma_rate.py  (on a B200): cycles per 128xNx16 MMA and the implied dense bf16 TFLOP/s/GPU."""
import ctypes, torch
lib = ctypes.CDLL("tools/experiments/mma_rate.so")
lib.mma_rate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
iters = 4096
for n, ctas in ((64, 148), (128, 148), (256, 148), (64, 296), (128, 296)):
    out = torch.zeros(ctas, dtype=torch.int64, device="cuda")
    rc = lib.mma_rate(out.data_ptr(), n, iters, ctas)
    cyc = out.float().mean().item()
    per_mma = cyc / (iters * 4)
    per_sm = per_mma / (ctas // 148)          # two co-resident CTAs share one tensor core
    flops = 2 * 128 * n * 16
    print("N=%3d ctas=%3d rc=%d  %.1f cycles/MMA per CTA  -> %.1f cycles/MMA per SM, %.0f TFLOP/s @1.9 GHz x 148 SMs"
          % (n, ctas, rc, per_mma, per_sm, flops / per_sm * 1.9e9 * 148 / 1e12))
