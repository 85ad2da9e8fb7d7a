"""Diagnostic timing of b200q.attn_bf16 at H=12, L=32760: full kernel vs tensor pipeline alone (+8) vs softmax alone (+16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q  # noqa: E402

H, L = 12, 32760
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(L, H * 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))


def timed(n=5):
    for _ in range(2):
        b200q.attn_bf16(q, k, v, H)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        b200q.attn_bf16(q, k, v, H)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


b200q.attn_bf16_set_fast(-1)
for rnd in range(2):
    for mode in (2, 66, 130, 26, 26 + 64, 26 + 128):
        b200q.attn_bf16_set_mode(mode)
        print(rnd, "mode", mode, f"{timed():.3f} ms", flush=True)
b200q.attn_bf16_set_mode(2)
for rnd in range(2):
    for fast in (3, 4):
      for wm in (0, 64, 128):
        b200q.attn_bf16_set_mode(2 + wm)
        b200q.attn_bf16_set_fast(fast)
        print(rnd, "max-free poly pairs", fast, "wait mode", wm, f"{timed():.3f} ms", flush=True)
b200q.attn_bf16_set_fast(3)
b200q.attn_bf16_set_mode(2)
