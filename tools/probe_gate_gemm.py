"""Gate-residual GEMM (fp32 residual stream in the epilogue) at the Wan shapes: time, TOP/s and HBM rate (A + W + residual in + out)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wan2.1-quantization_b200"))
import b200q
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=12):
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
    return ts[len(ts) // 2]


for (M, N, K) in [(32760, 1536, 1536), (4095, 1536, 1536), (32760, 1536, 8960), (75600, 5120, 5120)]:
    qa = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
    qw = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
    da = torch.rand(M, device=dev); dw = torch.rand(N, device=dev); zp = torch.ones(N, device=dev)
    rs = torch.ones(M, dtype=torch.int32, device=dev); bias = torch.rand(N, device=dev)
    res = torch.zeros(M, N, device=dev)
    o = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for mode in (0, 1, 3):
        b200q.gemm_set_cluster(mode)
        ms = timeit(lambda: b200q.gemm_w8a8(qa, qw, da, dw, zp, rs, bias, epilogue=b200q.EPI_GATE_RESIDUAL, residual=res, gate=dw))
        mb = (M * K + N * K + 8 * M * N) / 1e6
        print(f"gate {M}x{N}x{K} mode {mode}: {ms * 1e3:7.1f} us  {2 * M * N * K / ms / 1e9:7.0f} TOP/s  {mb / ms / 1e3:6.2f} TB/s of {mb:.0f} MB", flush=True)
    b200q.gemm_set_cluster(0)
    ms = timeit(lambda: b200q.gemm_w8a8(qa, qw, da, dw, zp, rs, bias, out=o))
    print(f"bf16 {M}x{N}x{K} auto  : {ms * 1e3:7.1f} us  {2 * M * N * K / ms / 1e9:7.0f} TOP/s", flush=True)
    del qa, qw, res, o
