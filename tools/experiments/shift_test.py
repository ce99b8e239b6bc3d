import ctypes, torch, subprocess, sys
lib = ctypes.CDLL("tools/experiments/shift_test.so")
lib.shift_test.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
v = int(sys.argv[1]); js = [int(x) for x in sys.argv[2:]]
torch.manual_seed(0)
A = torch.randn(256, 64, device="cuda").to(torch.bfloat16)
B = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(128, 64, device="cuda")
for j in js:
    rc = lib.shift_test(A.data_ptr(), B.data_ptr(), out.data_ptr(), j, v)
    ref = A[j:j + 128].float() @ B.float().t()
    err = float((out - ref).abs().max() / ref.abs().max())
    print("variant", v, "j", j, "rc", rc, "rel err %.3e" % err, flush=True)
