"""GPU probe: GEMM throughput per scheduling mode (1 single-CTA, 2 multicast pair, 3 cta_group::2 pair), raw ctypes loop."""
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


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [(32760, 1536, 1536), (32760, 4608, 1536), (32760, 8960, 1536), (32760, 1536, 8960), (75600, 5120, 5120),
          (75600, 13824, 5120), (8192, 8192, 8192)]
for (M, N, K) in shapes:
    qa = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
    w4 = b200q.pack_w4(torch.randint(-8, 8, (N, K), dtype=torch.int8, device=dev))
    da = torch.rand(M, device=dev); dw = torch.rand(N, device=dev); zp = torch.ones(N, device=dev)
    rs = torch.ones(M, dtype=torch.int32, device=dev); bias = torch.rand(N, device=dev)
    o = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    res = torch.zeros(M, N, device=dev) if N <= 5120 else None
    for mode in (1, 2, 3):
        b200q.gemm_set_cluster(mode)
        tag = f"{M}x{N}x{K}_mode{mode}"
        ms = timeit(lambda: b200q.gemm_w8a8(qa, qw, da, dw, zp, rs, bias, out=o))
        out["w8a8_" + tag] = 2 * M * N * K / ms / 1e9
        ms = timeit(lambda: b200q.gemm_w8a8(qa, qw, da, dw, zp, rs, bias, out=o, epilogue=b200q.EPI_GELU_TANH))
        out["gelu_" + tag] = 2 * M * N * K / ms / 1e9
        if res is not None:
            ms = timeit(lambda: b200q.gemm_w8a8(qa, qw, da, dw, zp, rs, bias, epilogue=b200q.EPI_GATE_RESIDUAL, residual=res, gate=dw))
            out["gate_" + tag] = 2 * M * N * K / ms / 1e9
        ms = timeit(lambda: b200q.gemm_w4a8(qa, w4, K, da, dw, zp, rs, bias, out=o))
        out["w4a8_" + tag] = 2 * M * N * K / ms / 1e9
    b200q.gemm_set_cluster(0)
    ms = timeit(lambda: torch._int_mm(qa, qw.t()))
    out[f"cublaslt_{M}x{N}x{K}"] = 2 * M * N * K / ms / 1e9
    del qa, qw, o, res, w4
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_gemm_modes.json"), "w"), indent=1)
