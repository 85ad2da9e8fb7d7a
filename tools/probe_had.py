"""Timing of b200q_had_quant_rows at the Wan shapes: warp-per-row register kernel vs shared-memory tile kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q
from qdiff.base.quant_layer import ActPlan
dev = torch.device("cuda")
lib = b200q.load()
for (M, n, dt) in [(32760, 1536, torch.float32), (75600, 5120, torch.float32), (75600, 5120, torch.bfloat16)]:
    x = torch.randn(M, n, device=dev).to(dt)
    plan = ActPlan.rotation(n, torch.ones(n), torch.rand(n) + 0.5, dev)
    q = torch.empty(M, n, dtype=torch.int8, device=dev)
    for mode in ((1, 2, 0) if n == 5120 else (1, 0)):
        lib.b200q_had_set_mode(mode)
        f = lambda: b200q.had_quant_rows(x, plan.colscale, plan.hadK, plan.K, plan.log2w, 8, out=q)
        for _ in range(3): f()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): f()
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) / 20 * 1e3
        by = M * n * (x.element_size() + 1) + 8 * M
        print(f"{M}x{n} {dt} K={plan.K} w=2^{plan.log2w} mode={('tile', 'warp', 'warp (3 CTAs/SM)')[mode]}: {us:.1f} us  {by / us / 1e3:.0f} GB/s  ({by / us / 1e3 / 6449.7:.2f} of HBM copy)", flush=True)
lib.b200q_had_set_mode(1)
