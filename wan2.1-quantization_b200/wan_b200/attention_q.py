"""Quantized attention (SURVEY §8 a-8 / row c), PARITY MODE: the reference's materialised fake-quant attention
(examples/Wan2.1/models/quant_opensora.py:430-478 with the 'row' attention-map group of
quant_utils/qdiff/base/quant_attn.py:168-174) executed with libb200q kernels:

  Q, K  per-(token, head) symmetric codes            b200q.quant_rows on [L*H, hd] rows                 (bit-exact)
  V     per-(head, channel) codes over all tokens    b200q.quant_rows on V^T [H*hd, L] rows             (bit-exact)
  S     = scale*dq[m]*dk[n] * (qq . kq^T)            tcgen05 int8 GEMM per head, int32 accumulators exact, fp32 out
  P     = softmax(S) (fp32), then ONE scale per key column over all queries: quant_rows on P^T rows
  O     = deq(P) @ deq(V)

The last product cannot be an integer GEMM: the attention-map scale delta_p[n] is indexed by the key n, i.e. by the
CONTRACTION index of P@V, so it does not factor out of the sum (a per-row/per-tile P scale would).  The reference has
the column maximum over all queries because it materialises P [H, L, L] (51 GB at L = 32,760 — it cannot run at the
BASELINE shapes); a fused one-pass kernel cannot.  This module is therefore the small-L parity path; the fused
kernel uses a per-tile P scale (fast mode) and is reported as such.
"""
from __future__ import annotations

import torch

import b200q


def quantized_attention_parity(q, k, v, num_heads, qk_bits=8, qk_sym=True, v_bits=8, v_sym=True, p_bits=8, p_sym=False,
                               quant_p=True, scale=None, return_info=False):
    """q [Lq, H*hd], k, v [Lk, H*hd] (fp32 / bf16, CUDA) -> [Lq, H*hd] fp32."""
    Lq, D = q.shape
    Lk = k.shape[0]
    H = num_heads
    hd = D // H
    scale = hd ** -0.5 if scale is None else scale
    if not (qk_sym and v_sym):
        raise NotImplementedError("asymmetric Q/K/V attention quantizers need zero-point cross terms; the spec uses sym")
    qq, dq, _, _ = b200q.quant_rows(q.reshape(Lq * H, hd), qk_bits, True, True, want_rowsum=False)
    kq, dk, _, _ = b200q.quant_rows(k.reshape(Lk * H, hd), qk_bits, True, True, want_rowsum=False)
    vq, dv, _, _ = b200q.quant_rows(v.t().contiguous(), v_bits, True, True, want_rowsum=False)        # [D, Lk], [D]
    qq, kq = qq.view(Lq, D), kq.view(Lk, D)
    dq, dk = dq.view(Lq, H), dk.view(Lk, H)
    out = torch.empty(Lq, D, dtype=torch.float32, device=q.device)
    info = dict(qq=qq, dq=dq, kq=kq, dk=dk, vq=vq, dv=dv, pq=[], dp=[], zp=[])
    for h in range(H):
        cols = slice(h * hd, (h + 1) * hd)
        da = (dq[:, h] * scale).contiguous()        # (q_dq * scale) @ k_dq^T  (quant_opensora.py:456-457)
        s = b200q.gemm_w8a8(qq[:, cols], kq[:, cols], da, dk[:, h].contiguous(), out_dtype=torch.float32)   # [Lq, Lk]
        p = torch.softmax(s, dim=-1)
        v_dq = b200q.dequant_rows(vq[cols], dv[cols], None)                                             # [hd, Lk]
        if quant_p:
            pt = p.t().contiguous()                                                                     # one row per key
            pq, dp, zp, _ = b200q.quant_rows(pt, p_bits, p_sym, True, want_rowsum=False)
            p_dq = b200q.dequant_rows(pq, dp, zp)                                                       # [Lk, Lq]
            out[:, cols] = p_dq.t() @ v_dq.t()
            if return_info:
                info["pq"].append(pq); info["dp"].append(dp); info["zp"].append(zp)
        else:
            out[:, cols] = p @ v_dq.t()
    return (out, info) if return_info else out
