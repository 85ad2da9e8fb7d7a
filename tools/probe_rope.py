"""rmsnorm_rope at [32760, 1536] (q slice of the fused q|k|v output): plain vs with the per-head norm maxima."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wan2.1-quantization_b200"))
import b200q
dev = torch.device("cuda:0")
L, D = 32760, 1536
qkv = torch.randn(L, 3 * D, device=dev).to(torch.bfloat16)
w = torch.rand(D, device=dev) + 0.5
cos = torch.rand(L, 64, device=dev); sin = torch.rand(L, 64, device=dev)
nrm = torch.zeros(12, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=15):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


mb = (2 * L * D * 2 + 2 * L * 64 * 4) / 1e6
for name, fn in (("plain", lambda: b200q.rmsnorm_rope(qkv[:, :D], w, 1e-6, cos, sin, 128)),
                 ("with head norm maxima", lambda: b200q.rmsnorm_rope(qkv[:, :D], w, 1e-6, cos, sin, 128, head_sq_max=nrm)),
                 ("no rope, with maxima", lambda: b200q.rmsnorm_rope(qkv[:, :D], w, 1e-6, None, None, 0, head_sq_max=nrm))):
    us = timeit(fn)
    print(f"rmsnorm_rope {name:24s} {us:6.1f} us  {mb / us / 1e3:5.2f} TB/s of {mb:.0f} MB")
