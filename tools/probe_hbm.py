"""GPU probe of the HBM-bound kernels without Python-wrapper overhead in the timed region: outputs pre-allocated,
raw ctypes entry points called back-to-back (inputs larger than L2, so no flush), CUDA events around 20 launches."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wan2.1-quantization_b200"))
import b200q  # noqa: E402

lib = b200q.load()
dev = torch.device("cuda:0")
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
DT = {torch.float32: 0, torch.bfloat16: 1}
out = {}


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for (M, K, dt) in [(32760, 1536, torch.float32), (32760, 1536, torch.bfloat16), (32760, 8960, torch.bfloat16),
                   (75600, 5120, torch.float32), (75600, 13824, torch.bfloat16)]:
    tag = f"{M}x{K}_{str(dt)[6:]}"
    x = torch.randn(M, K, device=dev, dtype=dt)
    q = torch.empty(M, K, dtype=torch.int8, device=dev)
    d = torch.empty(M, device=dev); z = torch.empty(M, device=dev); rs = torch.empty(M, dtype=torch.int32, device=dev)
    stat = torch.zeros(K, device=dev); sh = torch.randn(K, device=dev) * 0.1
    es = x.element_size()
    ms = timeit(lambda: lib.b200q_quant_rows(P(x), DT[dt], M, K, K, 8, 1, 1, P(q), K, P(d), P(z), P(rs), None, None, st))
    out[f"quant_{tag}_gbs"] = (M * K * (es + 1) + 12 * M) / ms / 1e6
    ms = timeit(lambda: lib.b200q_calib_absmax_minmax(P(x), DT[dt], M, K, K, P(stat), None, None, st))
    out[f"calib_{tag}_gbs"] = (M * K * es + 12 * K) / ms / 1e6
    ms = timeit(lambda: lib.b200q_ln_mod_quant(P(x), DT[dt], M, K, K, None, None, ctypes.c_float(1e-6), P(sh), P(sh), 8,
                                               P(q), K, P(d), P(rs), None, 0, 0, st))
    out[f"lnq_{tag}_gbs"] = (M * K * (es + 1) + 8 * M) / ms / 1e6
    if dt == torch.bfloat16 and K % 128 == 0:
        o = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
        cs = torch.rand(M, 64, device=dev)
        ms = timeit(lambda: lib.b200q_rmsnorm_rope(P(x), 1, M, K, K, P(sh), ctypes.c_float(1e-6), P(cs), P(cs), 128, P(o), K, st))
        out[f"rmsrope_{tag}_gbs"] = (M * K * 4 + M * 512) / ms / 1e6
        del o
    y = torch.empty_like(x)
    ms = timeit(lambda: y.copy_(x))
    out[f"copy_{tag}_gbs"] = 2 * M * K * es / ms / 1e6
    del x, q, y
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_hbm.json"), "w"), indent=1)
