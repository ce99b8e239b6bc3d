import ctypes, torch, sys
lib = ctypes.CDLL("tools/experiments/noswz_test.so")
lib.noswz_test.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
torch.manual_seed(0)
buf = torch.randn(256, 8, device="cuda").to(torch.bfloat16)
B = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(128, 64, device="cuda")
for v in (0, 1):
    for shift in (0, 1, 5, 8):
        out.zero_()
        rc = lib.noswz_test(buf.data_ptr(), B.data_ptr(), out.data_ptr(), shift, v)
        # A'[m][8c + e] = buf[m + c + shift][e]
        A = torch.stack([buf[shift + c: shift + c + 128] for c in range(8)], 1).reshape(128, 64).float()
        ref = A @ B.float().t()
        err = float((out - ref).abs().max() / ref.abs().max())
        print("variant", v, "shift", shift, "rc", rc, "rel err %.3e" % err, flush=True)
