"""One-off GPU probe: int8 cuBLASLt peak (torch._int_mm), and first timings of the b200q kernels.
Writes gpurun_out/probe.json.  Not part of the product or the bench contract."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wan2.1-quantization_b200"))
import b200q  # noqa: E402

dev = torch.device("cuda:0")
out = {}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3, flush_l2=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


# int8 peak via cuBLASLt
for n in (8192,):
    a = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev).t()
    try:
        med, best = timeit(lambda: torch._int_mm(a, b), iters=10, flush_l2=False)
        out[f"int_mm_{n}_tops_best"] = 2 * n ** 3 / best / 1e9
        out[f"int_mm_{n}_tops_med"] = 2 * n ** 3 / med / 1e9
    except Exception as ex:  # noqa: BLE001
        out["int_mm_error"] = repr(ex)
    a16 = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    med, best = timeit(lambda: a16 @ a16, iters=10, flush_l2=False)
    out[f"bf16_{n}_tflops_best"] = 2 * n ** 3 / best / 1e9

for (M, N, K) in [(32760, 1536, 1536), (32760, 8960, 1536), (32760, 1536, 8960), (8192, 8192, 8192),
                  (75600, 5120, 5120)]:
    qa = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
    da = torch.rand(M, device=dev); dw = torch.rand(N, device=dev); zp = torch.ones(N, device=dev)
    rs = torch.ones(M, dtype=torch.int32, device=dev); bias = torch.rand(N, device=dev)
    o = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    try:
        med, best = timeit(lambda: b200q.gemm_w8a8(qa, qw, da, dw, zp, rs, bias, out=o), iters=10)
        out[f"gemm_{M}x{N}x{K}_tops_med"] = 2 * M * N * K / med / 1e9
        out[f"gemm_{M}x{N}x{K}_tops_best"] = 2 * M * N * K / best / 1e9
        med, best = timeit(lambda: torch._int_mm(qa, qw.t()), iters=10)
        out[f"int_mm_{M}x{N}x{K}_tops_med"] = 2 * M * N * K / med / 1e9
    except Exception as ex:  # noqa: BLE001
        out[f"gemm_{M}x{N}x{K}_error"] = repr(ex)
    del qa, qw, o

for (M, K, dt) in [(32760, 1536, torch.float32), (32760, 1536, torch.bfloat16), (32760, 8960, torch.bfloat16),
                   (75600, 5120, torch.float32), (75600, 13824, torch.bfloat16)]:
    x = torch.randn(M, K, device=dev, dtype=dt)
    q = torch.empty(M, K, dtype=torch.int8, device=dev)
    med, best = timeit(lambda: b200q.quant_rows(x, 8, True, True, out=q))
    nbytes = M * K * (x.element_size() + 1) + 12 * M
    out[f"quant_{M}x{K}_{str(dt)[6:]}_gbs_med"] = nbytes / med / 1e6
    out[f"quant_{M}x{K}_{str(dt)[6:]}_gbs_best"] = nbytes / best / 1e6
    a = torch.zeros(K, device=dev)
    med, best = timeit(lambda: b200q.calib_update(x, a))
    out[f"calib_{M}x{K}_{str(dt)[6:]}_gbs_med"] = (M * K * x.element_size() + 12 * K) / med / 1e6
    if dt == torch.float32:
        sh = torch.randn(K, device=dev)
        med, best = timeit(lambda: b200q.ln_mod_quant(x, 1e-6, None, None, sh, sh, 8))
        out[f"lnq_{M}x{K}_gbs_med"] = nbytes / med / 1e6
    y = torch.empty_like(x)
    med, best = timeit(lambda: y.copy_(x))
    out[f"copy_{M}x{K}_{str(dt)[6:]}_gbs_med"] = 2 * M * K * x.element_size() / med / 1e6
    del x, q, y

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
