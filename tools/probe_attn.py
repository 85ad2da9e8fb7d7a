"""GPU probe: which library attention backend is fastest for the Wan shapes (not part of the product)."""
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

dev = "cuda"
for (H, L, Lk) in [(12, 32760, 32760), (12, 32760, 512)]:
    q = torch.randn(1, H, L, 128, device=dev, dtype=torch.bfloat16)
    k = torch.randn(1, H, Lk, 128, device=dev, dtype=torch.bfloat16)
    v = torch.randn(1, H, Lk, 128, device=dev, dtype=torch.bfloat16)
    # token-major views like the runtime produces
    qs = torch.randn(L, H, 128, device=dev, dtype=torch.bfloat16).permute(1, 0, 2).unsqueeze(0)
    ks = torch.randn(Lk, H, 128, device=dev, dtype=torch.bfloat16).permute(1, 0, 2).unsqueeze(0)
    vs = torch.randn(Lk, H, 128, device=dev, dtype=torch.bfloat16).permute(1, 0, 2).unsqueeze(0)
    for name, be in [("default", None), ("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION)]:
        for tag, (a, b, c) in (("contig", (q, k, v)), ("strided", (qs, ks, vs))):
            try:
                def run():
                    if be is None:
                        return F.scaled_dot_product_attention(a, b, c)
                    with sdpa_kernel(be):
                        return F.scaled_dot_product_attention(a, b, c)
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(5):
                    run()
                e.record(); torch.cuda.synchronize()
                ms = s.elapsed_time(e) / 5
                print(f"H={H} L={L} Lk={Lk} {name:9s} {tag:8s} {ms:8.3f} ms  {4 * H * L * Lk * 128 / ms / 1e9:8.1f} TFLOP/s", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"H={H} L={L} Lk={Lk} {name} {tag} failed: {str(ex)[:100]}", flush=True)
try:
    from flash_attn import flash_attn_func
    q = torch.randn(1, 32760, 12, 128, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        flash_attn_func(q, q, q)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        flash_attn_func(q, q, q)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f"flash_attn pkg {ms:.3f} ms {4 * 12 * 32760 ** 2 * 128 / ms / 1e9:.1f} TFLOP/s")
except Exception as ex:  # noqa: BLE001
    print("flash_attn pkg failed:", str(ex)[:200])
