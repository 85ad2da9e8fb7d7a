"""Time b200q_attn_i8 (fused int8 attention) against the library bf16 flash attention at the Wan shapes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wan2.1-quantization_b200")]
import b200q  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    dev = torch.device("cuda:0")
    res = {}
    shapes = [(12, 32760, 32760), (12, 32760, 512), (3, 32760, 32760), (40, 9450, 9450)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    for H, Lq, Lk in shapes:
        D = H * 128
        g = torch.Generator(device=dev).manual_seed(0)
        q = torch.randn(Lq, D, device=dev, generator=g, dtype=torch.bfloat16)
        k = torch.randn(Lk, D, device=dev, generator=g, dtype=torch.bfloat16)
        v = torch.randn(Lk, D, device=dev, generator=g, dtype=torch.bfloat16)
        qq, dq, _, _ = b200q.quant_rows(q.view(Lq * H, 128), 8, True, True, want_rowsum=False)
        kq, dk, _, _ = b200q.quant_rows(k.view(Lk * H, 128), 8, True, True, want_rowsum=False)
        vt, dv = b200q.quant_vt(v, 8)
        qq, kq, dq, dk = qq.view(Lq, D), kq.view(Lk, D), dq.view(Lq, H), dk.view(Lk, H)
        out = torch.empty(Lq, D, dtype=torch.bfloat16, device=dev)
        modes = {}
        for mode in (2, 6, 10, 14):
            b200q.load().b200q_attn_set_mode(mode)
            modes[mode] = timed(lambda: b200q.attn_i8(qq, dq, kq, dk, vt, dv, H, out=out))
        best = min(modes, key=modes.get)
        b200q.load().b200q_attn_set_mode(best)
        ms = timed(lambda: b200q.attn_i8(qq, dq, kq, dk, vt, dv, H, out=out))
        flops = 4.0 * Lq * Lk * D
        qh = q.view(Lq, H, 128).permute(1, 0, 2).unsqueeze(0)
        kh = k.view(Lk, H, 128).permute(1, 0, 2).unsqueeze(0)
        vh = v.view(Lk, H, 128).permute(1, 0, 2).unsqueeze(0)
        ms_lib = timed(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
        ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)[0].permute(1, 0, 2).reshape(Lq, D).float()
        o = out.float()
        cos = float((o.flatten() @ ref.flatten()) / (o.norm() * ref.norm()))
        ms_qv = timed(lambda: b200q.quant_vt(v, 8))
        ms_qq = timed(lambda: b200q.quant_rows(q.view(Lq * H, 128), 8, True, True, want_rowsum=False))
        key = f"H{H}_Lq{Lq}_Lk{Lk}"
        res[key] = {"attn_i8_ms": ms, "attn_i8_tflops_equiv": flops / ms / 1e9, "int8_tops_3pass": 1.5 * flops / ms / 1e9,
                    "sdpa_bf16_ms": ms_lib, "sdpa_tflops": flops / ms_lib / 1e9, "cos_vs_bf16_sdpa": cos,
                    "quant_vt_ms": ms_qv, "modes_ms": modes, "quant_rows_q_ms": ms_qq}
        print(key, json.dumps(res[key]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_attn_i8.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
