"""One b200q.attn_bf16 launch per iteration at a given shape (ncu target).  python tools/run_attn_bf16.py H Lq Lk [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q  # noqa: E402

H, Lq, Lk = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(L, H * 128, device="cuda", generator=g).to(torch.bfloat16) for L in (Lq, Lk, Lk))
for _ in range(iters):
    o = b200q.attn_bf16(q, k, v, H)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
