import ctypes, os, sys, torch
sys.path.insert(0, "/root/repo/wan2.1-quantization_b200")
import b200q
lib = b200q.load(); dev = torch.device("cuda:0")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, K) in [(32760, 1536), (75600, 5120)]:
    x = torch.randn(M, K, device=dev); q = torch.empty(M, K, dtype=torch.int8, device=dev)
    d = torch.empty(M, device=dev); rs = torch.empty(M, dtype=torch.int32, device=dev); sh = torch.randn(K, device=dev) * 0.1
    fn = lambda: lib.b200q_ln_mod_quant(P(x), 0, M, K, K, None, None, ctypes.c_float(1e-6), P(sh), P(sh), 8, P(q), K, P(d), P(rs), None, 0, 0, st)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(os.environ.get("B200Q_LN_VMAX"), M, K, f"{(M*K*5+8*M)/ms/1e6:.0f} GB/s")
