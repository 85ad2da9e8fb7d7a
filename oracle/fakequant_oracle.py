"""CPU ORACLE (test infrastructure, NOT product code).

A plain-PyTorch, CPU, fp32 restatement of the reference's fake-quant arithmetic
for the hot path of BASELINE.json's north_star.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; nothing under `wan2.1-quantization_b200/` does.

Parity pinning: every function here is checked against the *imported, unmodified*
reference code (`/root/reference/ViDiT-Q/quant_utils/qdiff`) by
`tests/test_oracle_vs_reference.py` (runs where /root/reference exists) and
against the golden vectors that `oracle/gen_golden.py` produced from that imported
reference (`tests/golden/*.pt`, run everywhere).

All citations are relative to /root/reference/ViDiT-Q/.
dtype contract (SURVEY §7 hard-part 1): inputs are up-cast to fp32 and all
quantizer arithmetic is IEEE fp32: '/' is a true division, round is
round-half-to-even (torch.round).
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# quantizers
# --------------------------------------------------------------------------
def n_levels_of(n_bits: int, sym: bool) -> int:
    """quant_utils/qdiff/base/base_quantizer.py:32"""
    return 2 ** (n_bits - 1) - 1 if sym else 2 ** n_bits


def quant_params_rows(x: torch.Tensor, n_bits: int, sym: bool, dynamic: bool):
    """Per-row (delta, zero_point), both [G,1] fp32.

    dynamic=True  -> DynamicQuantizer.quantize   base_quantizer.py:110-151
    dynamic=False -> StaticQuantizer.init_quant_params  base_quantizer.py:70-99
    The only numeric difference is the 1e-6 floor on delta that the dynamic
    symmetric path applies (`delta[delta < eps] = eps`, :122-128).
    """
    assert x.dim() == 2
    x = x.float()
    nl = n_levels_of(n_bits, sym)
    if sym:
        amax = x.abs().max(dim=1)[0]
        delta = amax / nl
        zp = torch.zeros_like(delta)
        if dynamic:
            delta = torch.where(delta < 1.0e-6, torch.full_like(delta, 1.0e-6), delta)
    else:
        xmax = x.max(dim=1)[0].clamp_min(0.0)
        xmin = x.min(dim=1)[0].clamp_max(0.0)
        delta = (xmax - xmin) / (nl - 1)
        zp = torch.round(xmin / delta) + (nl / 2)
    return delta.unsqueeze(-1), zp.unsqueeze(-1)


def quant_codes_rows(x: torch.Tensor, delta: torch.Tensor, zp: torch.Tensor, n_bits: int, sym: bool):
    """x_quant of base_quantizer.py:155-157 / :64-68 (float tensor of integers)."""
    nl = n_levels_of(n_bits, sym)
    x_int = torch.round(x.float() / delta) - zp
    return torch.clamp(x_int, -nl - 1, nl)


def quant_rows(x: torch.Tensor, n_bits: int = 8, sym: bool = True, dynamic: bool = True):
    """-> (codes float [G,C], delta [G,1], zero_point [G,1])"""
    delta, zp = quant_params_rows(x, n_bits, sym, dynamic)
    return quant_codes_rows(x, delta, zp, n_bits, sym), delta, zp


def dequant_rows(q: torch.Tensor, delta: torch.Tensor, zp: torch.Tensor):
    """base_quantizer.py:159-162: (x_quant + zero_point) * delta"""
    return (q + zp) * delta


def fake_quant_rows(x, n_bits=8, sym=True, dynamic=True):
    q, d, z = quant_rows(x, n_bits, sym, dynamic)
    return dequant_rows(q, d, z)


def forward_with_quant_params(x, delta, n_bits=8):
    """DynamicQuantizer.forward_with_quant_params (mixed_precision=None branch),
    base_quantizer.py:164-206: unsigned codes in [0, 2*n_levels+1]."""
    nl = n_levels_of(n_bits, True)
    delta = torch.where(delta < 1.0e-6, torch.full_like(delta, 1.0e-6), delta)
    delta = delta / (nl * 2 + 1)
    q = torch.clamp(torch.round(x / delta), 0, nl * 2 + 1)
    return q * delta


# --------------------------------------------------------------------------
# quantized linear
# --------------------------------------------------------------------------
def quantized_linear_fake(x, weight, bias, w_bits=8, w_sym=False, a_bits=8, a_sym=True):
    """QuantizedLinear.forward, quant_layer.py:57-74, with the weight fake-quantised
    as in __init__ (:40).  x: [B,N,C] or [M,C]."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).float()
    w_dq = fake_quant_rows(weight.float(), w_bits, w_sym, dynamic=False)
    x_dq = fake_quant_rows(x2, a_bits, a_sym, dynamic=True)
    y = F.linear(x_dq, w_dq, None if bias is None else bias.float())
    return y.reshape(*shp[:-1], weight.shape[0])


def int_accumulators(qa: torch.Tensor, qw: torch.Tensor) -> torch.Tensor:
    """Exact int32 accumulators  acc[m,n] = sum_k qa[m,k]*qw[n,k]  (SURVEY §8 a-5)."""
    return (qa.to(torch.int32) @ qw.to(torch.int32).t()).to(torch.int32)


def quantized_linear_int(x, weight, bias, w_bits=8, w_sym=False, a_bits=8, a_sym=True):
    """Algebraic real-integer form of quantized_linear_fake:
         y[m,n] = da[m]*dw[n]*( acc[m,n] + zp_w[n]*rowsum_a[m] + zp_a[m]*colsum_w[n]
                                + K*zp_a[m]*zp_w[n] ) + bias[n]
    returns (y fp32, acc int32, qa, da, zpa, qw, dw, zpw)."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1]).float()
    qa, da, zpa = quant_rows(x2, a_bits, a_sym, True)
    qw, dw, zpw = quant_rows(weight.float(), w_bits, w_sym, False)
    acc = int_accumulators(qa, qw)
    K = x2.shape[1]
    rs = qa.sum(dim=1, keepdim=True)            # [M,1]
    cs = qw.sum(dim=1, keepdim=True).t()        # [1,N]
    full = acc.double() + zpw.t().double() * rs.double() + zpa.double() * cs.double() \
        + K * zpa.double() * zpw.t().double()
    y = (da.double() * dw.t().double() * full)
    if bias is not None:
        y = y + bias.double()
    return y.float().reshape(*shp[:-1], weight.shape[0]), acc, qa, da, zpa, qw, dw, zpw


# --------------------------------------------------------------------------
# calibration
# --------------------------------------------------------------------------
def calib_absmax(x: torch.Tensor) -> torch.Tensor:
    """SaveActivationHook.__call__ default branch,
    examples/Wan2.1/get_calib_data_wanx.py:262-263."""
    C = x.shape[-1]
    return x.reshape(-1, C).abs().max(dim=0)[0]


def calib_merge(per_call_stats) -> torch.Tensor:
    """gather_and_save_activation stack (:448) + rank-0 cat (:468) followed by the
    consumer's `.max(dim=0)[0]` (examples/Wan2.1/ptq_wanx.py:336) == running max."""
    return torch.stack(list(per_call_stats), dim=0).max(dim=0)[0]


def calib_minmax(x: torch.Tensor):
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    return x2.min(dim=0)[0], x2.max(dim=0)[0]


# --------------------------------------------------------------------------
# quantized attention (spec: OpenSORA adapter)
# --------------------------------------------------------------------------
def quantized_attention_fake(q, k, v, qk_bits=8, qk_sym=True, v_bits=8, v_sym=True,
                             p_bits=8, p_sym=False, quant_p=True, scale=None):
    """examples/Wan2.1/models/quant_opensora.py:430-478 with the 'row' attn-map group
    of quant_utils/qdiff/base/quant_attn.py:168-174.
    q,k,v: [B,H,L,hd] fp32.  Q,K per-(token,head); V per-(head,channel) over tokens;
    P per key column over all queries.  Returns (out [B,H,Lq,hd], dict of codes)."""
    B, H, Lq, hd = q.shape
    Lk = k.shape[2]
    scale = hd ** -0.5 if scale is None else scale
    qq, dq, zq = quant_rows(q.reshape(-1, hd), qk_bits, qk_sym, True)
    kq, dk, zk = quant_rows(k.reshape(-1, hd), qk_bits, qk_sym, True)
    vq, dv, zv = quant_rows(v.permute(0, 1, 3, 2).reshape(-1, Lk), v_bits, v_sym, True)
    q_dq = dequant_rows(qq, dq, zq).reshape(B, H, Lq, hd)
    k_dq = dequant_rows(kq, dk, zk).reshape(B, H, Lk, hd)
    v_dq = dequant_rows(vq, dv, zv).reshape(B, H, hd, Lk).permute(0, 1, 3, 2)
    attn = ((q_dq * scale) @ k_dq.transpose(-2, -1)).float().softmax(dim=-1)
    info = dict(qq=qq.reshape(B, H, Lq, hd), dq=dq.reshape(B, H, Lq), kq=kq.reshape(B, H, Lk, hd),
                dk=dk.reshape(B, H, Lk), vq=vq.reshape(B, H, hd, Lk), dv=dv.reshape(B, H, hd),
                zq=zq.reshape(B, H, Lq), zk=zk.reshape(B, H, Lk), zv=zv.reshape(B, H, hd))
    if quant_p:
        pt = attn.permute(0, 1, 3, 2).reshape(-1, Lq)           # one row per key column
        pq, dp, zp = quant_rows(pt, p_bits, p_sym, True)
        attn = dequant_rows(pq, dp, zp).reshape(B, H, Lk, Lq).permute(0, 1, 3, 2)
        info.update(pq=pq.reshape(B, H, Lk, Lq).permute(0, 1, 3, 2), dp=dp.reshape(B, H, Lk),
                    zp=zp.reshape(B, H, Lk))
    return attn @ v_dq, info


def quantized_attention_rowstep(q, k, v, qk_bits=8, v_bits=8, p_bits=8, scale=None):
    """Fused-kernel ("fast mode") attention semantics, restated with the reference's own quantizers:
    Q,K per-(token,head) and V per-(head,channel) DynamicQuantizers exactly as quantized_attention_fake
    (quant_opensora.py:430-442); S = (q_dq*scale) @ k_dq^T, fp32 softmax (:456-462); the attention map is quantized by
    DynamicQuantizer.forward_with_quant_params (base_quantizer.py:164-206, unsigned [0, 2^b-1] grid) with
    delta = the row maximum of P (one step per query row) instead of QuantizedAttentionMapOpenSORA's per-key-column
    grouping.  Returns (out [B,H,Lq,hd], info) with info['p_codes'] = round(P / (rowmax/255)) in [0,255]."""
    B, H, Lq, hd = q.shape
    Lk = k.shape[2]
    scale = hd ** -0.5 if scale is None else scale
    qq, dq, zq = quant_rows(q.reshape(-1, hd), qk_bits, True, True)
    kq, dk, zk = quant_rows(k.reshape(-1, hd), qk_bits, True, True)
    vq, dv, zv = quant_rows(v.permute(0, 1, 3, 2).reshape(-1, Lk), v_bits, True, True)
    q_dq = dequant_rows(qq, dq, zq).reshape(B, H, Lq, hd)
    k_dq = dequant_rows(kq, dk, zk).reshape(B, H, Lk, hd)
    v_dq = dequant_rows(vq, dv, zv).reshape(B, H, hd, Lk).permute(0, 1, 3, 2)
    attn = ((q_dq * scale) @ k_dq.transpose(-2, -1)).float().softmax(dim=-1)
    pmax = attn.max(dim=-1, keepdim=True)[0].expand_as(attn).clone()
    p_dq = forward_with_quant_params(attn, pmax, p_bits)
    nl = n_levels_of(p_bits, True) * 2 + 1
    step = torch.where(pmax < 1.0e-6, torch.full_like(pmax, 1.0e-6), pmax) / nl
    info = dict(qq=qq.reshape(B, H, Lq, hd), dq=dq.reshape(B, H, Lq), kq=kq.reshape(B, H, Lk, hd),
                dk=dk.reshape(B, H, Lk), vq=vq.reshape(B, H, hd, Lk), dv=dv.reshape(B, H, hd), attn=attn,
                p_codes=torch.clamp(torch.round(attn / step), 0, nl), p_dequant=p_dq)
    return p_dq @ v_dq, info


# --------------------------------------------------------------------------
# one Wan DiT block (restatement; the reference block cannot run on CPU)
# --------------------------------------------------------------------------
def rope_params(max_seq_len, dim, theta=10000):
    """examples/Wan2.1/wan/modules/model.py:31-40"""
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def wan_freqs(head_dim):
    """model.py:527-533"""
    d = head_dim
    return torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                      rope_params(1024, 2 * (d // 6))], dim=1)


def rope_apply(x, grid, freqs, pos_offset=0):
    """model.py:43-70 for one sample; x [L,n,d]; grid=(f,h,w); float64 complex.
    pos_offset/L slice = the rank-sliced variant of
    wan/distributed/xdit_context_parallel.py:52-58."""
    L, n, d2 = x.shape[0], x.shape[1], x.shape[2] // 2
    f, h, w = grid
    fr = freqs.split([d2 - 2 * (d2 // 3), d2 // 3, d2 // 3], dim=1)
    fi = torch.cat([fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
                    fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                    fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, 1, -1)
    fi = fi[pos_offset:pos_offset + L]
    xc = torch.view_as_complex(x.to(torch.float64).reshape(L, n, -1, 2))
    return torch.view_as_real(xc * fi).flatten(2).float()


def rms_norm(x, weight, eps):
    """WanRMSNorm, model.py:73-89"""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * weight


def layer_norm(x, weight, bias, eps):
    """WanLayerNorm, model.py:92-102"""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).type_as(x)


class WanBlockOracle:
    """WanAttentionBlock.forward (model.py:293-370) with self-attention as intended by
    wan/distributed/xdit_context_parallel.py:155-192 at sp_size=1 (the in-file version
    has q un-normed/un-viewed, model.py:145-146, and raises).  Every nn.Linear is the
    fake-quant QuantizedLinear (quant_layer.py:57-74).  fp32 on CPU, batch 1.

    params: dict name -> tensor with keys
      self_attn.{q,k,v,o}.{weight,bias}, self_attn.norm_{q,k}.weight,
      cross_attn.{q,k,v,o}.{weight,bias}, cross_attn.norm_{q,k}.weight,
      norm3.{weight,bias}, ffn.0.{weight,bias}, ffn.2.{weight,bias}, modulation [1,6,D]
    """

    def __init__(self, params, dim, ffn_dim, num_heads, eps=1e-6, w_bits=8, w_sym=False,
                 a_bits=8, a_sym=True, quant=True, attn_quant=None, w_bits_by_layer=None):
        self.p = {k: v.float() for k, v in params.items()}
        self.dim, self.ffn_dim, self.h, self.eps = dim, ffn_dim, num_heads, eps
        self.qa = dict(w_bits=w_bits, w_sym=w_sym, a_bits=a_bits, a_sym=a_sym)
        self.quant = quant
        self.attn_quant = attn_quant      # None or dict(kwargs of quantized_attention_fake)
        self.w_bits_by_layer = w_bits_by_layer or {}
        self.freqs = wan_freqs(dim // num_heads)

    def lin(self, name, x):
        w, b = self.p[name + ".weight"], self.p.get(name + ".bias")
        if not self.quant:
            return F.linear(x, w, b)
        kw = dict(self.qa)
        if name in self.w_bits_by_layer:
            kw["w_bits"] = self.w_bits_by_layer[name]
        return quantized_linear_fake(x, w, b, **kw)

    def cross_attention(self, q, k, v):
        return self._attend(q, k, v)

    def attention(self, q, k, v):
        return self._attend(q, k, v)

    def _attend(self, q, k, v):
        # q,k,v [L,n,d] -> [1,n,L,d]
        q, k, v = (t.permute(1, 0, 2).unsqueeze(0) for t in (q, k, v))
        if self.attn_quant is not None and self.attn_quant.get("mode") == "rowstep":
            o, _ = quantized_attention_rowstep(q, k, v, **{a: b for a, b in self.attn_quant.items() if a != "mode"})
        elif self.attn_quant is not None:
            o, _ = quantized_attention_fake(q, k, v, **self.attn_quant)
        else:
            o = F.scaled_dot_product_attention(q, k, v)     # wan/modules/attention.py:171-178
        return o.squeeze(0).permute(1, 0, 2).flatten(1)

    def forward(self, x, e, grid, context, pos_offset=0, kv_full=None):
        """x [L,D] fp32, e [6,D] fp32, context [T,D]."""
        p, n, d = self.p, self.h, self.dim // self.h
        e = (p["modulation"].reshape(6, -1) + e).unbind(0)
        L = x.shape[0]
        h = layer_norm(x, None, None, self.eps) * (1 + e[1]) + e[0]
        q = rms_norm(self.lin("self_attn.q", h), p["self_attn.norm_q.weight"], self.eps).view(L, n, d)
        k = rms_norm(self.lin("self_attn.k", h), p["self_attn.norm_k.weight"], self.eps).view(L, n, d)
        v = self.lin("self_attn.v", h).view(L, n, d)
        q = rope_apply(q, grid, self.freqs, pos_offset)
        k = rope_apply(k, grid, self.freqs, pos_offset)
        y = self.lin("self_attn.o", self.attention(q, k, v))
        x = x + y * e[2]
        # cross attention  model.py:180-200
        h = layer_norm(x, p.get("norm3.weight"), p.get("norm3.bias"), self.eps)
        T = context.shape[0]
        q = rms_norm(self.lin("cross_attn.q", h), p["cross_attn.norm_q.weight"], self.eps).view(L, n, d)
        k = rms_norm(self.lin("cross_attn.k", context), p["cross_attn.norm_k.weight"], self.eps).view(T, n, d)
        v = self.lin("cross_attn.v", context).view(T, n, d)
        o = self.cross_attention(q, k, v)
        x = x + self.lin("cross_attn.o", o)
        # ffn  model.py:286-288, :359-362
        h = layer_norm(x, None, None, self.eps) * (1 + e[4]) + e[3]
        h = F.gelu(self.lin("ffn.0", h), approximate="tanh")
        y = self.lin("ffn.2", h)
        return x + y * e[5]


def make_block_params(dim, ffn_dim, seed=0, bias_std=0.02):
    """WanModel.init_weights (model.py:658-680): xavier-uniform linears; biases
    re-drawn N(0,bias_std) (SURVEY §8d config 1 — zero biases would hide bias bugs)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    for att in ("self_attn", "cross_attn"):
        for l in "qkvo":
            p[f"{att}.{l}.weight"] = xavier(dim, dim)
            p[f"{att}.{l}.bias"] = torch.randn(dim, generator=g) * bias_std
        p[f"{att}.norm_q.weight"] = torch.ones(dim)
        p[f"{att}.norm_k.weight"] = torch.ones(dim)
    p["norm3.weight"] = torch.ones(dim)
    p["norm3.bias"] = torch.zeros(dim)
    p["ffn.0.weight"] = xavier(ffn_dim, dim)
    p["ffn.0.bias"] = torch.randn(ffn_dim, generator=g) * bias_std
    p["ffn.2.weight"] = xavier(dim, ffn_dim)
    p["ffn.2.bias"] = torch.randn(dim, generator=g) * bias_std
    p["modulation"] = torch.randn(1, 6, dim, generator=g) / dim ** 0.5
    return p


# --------------------------------------------------------------------------
# full DiT step
# --------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim, position):
    """model.py:17-28 (float64)"""
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


class WanDiTOracle:
    """WanModel.forward (wan/modules/model.py:539-631) for ONE sample, fp32 on CPU, restated over a state dict with
    WanModel's parameter names: patch embedding (Conv3d with stride == kernel, :508-509), sinusoidal time embedding ->
    time_embedding -> time_projection (:585-590), text_embedding on the zero-padded context (:593-599), the blocks
    (WanBlockOracle), Head (:388-401: LayerNorm, modulation, linear) and unpatchify (:633-656).
    `lin(name, x, default)` (optional) overrides individual block linears, name = 'blocks.<i>.<layer>' — used by the
    tests to model layers the quant config keeps FP or runs as SmoothQuant / QuaRot / ViDiT-Q variants."""

    def __init__(self, sd, dim, ffn_dim, num_heads, num_layers, freq_dim=256, text_len=512, patch_size=(1, 2, 2),
                 out_dim=16, eps=1e-6, lin=None, **block_kw):
        self.sd = {k: v.detach().float().cpu() for k, v in sd.items() if torch.is_tensor(v)}
        self.dim, self.freq_dim, self.text_len, self.patch_size, self.out_dim, self.eps = dim, freq_dim, text_len, patch_size, out_dim, eps
        self.blocks = []
        for i in range(num_layers):
            pre = f"blocks.{i}."
            p = {k[len(pre):]: v for k, v in self.sd.items() if k.startswith(pre)}
            blk = WanBlockOracle(p, dim, ffn_dim, num_heads, eps=eps, **block_kw)
            if lin is not None:
                default = blk.lin
                blk.lin = (lambda name, x, _pre=pre, _d=default: lin(_pre + name, x, lambda: _d(name, x)))
            self.blocks.append(blk)

    def embed(self, latent, t, context):
        """-> x [L, D], e [1, D], e0 [6, D], ctx [text_len, D], grid"""
        sd = self.sd
        x = F.conv3d(latent.float().unsqueeze(0), sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=self.patch_size)
        grid = tuple(x.shape[2:])
        x = x.flatten(2).transpose(1, 2)[0]
        te = sinusoidal_embedding_1d(self.freq_dim, t.reshape(-1)[:1]).float()
        e = F.linear(F.silu(F.linear(te, sd["time_embedding.0.weight"], sd["time_embedding.0.bias"])),
                     sd["time_embedding.2.weight"], sd["time_embedding.2.bias"])
        e0 = F.linear(F.silu(e), sd["time_projection.1.weight"], sd["time_projection.1.bias"]).view(6, self.dim)
        ctx = torch.zeros(self.text_len, context.shape[1])
        ctx[:context.shape[0]] = context.float()
        ctx = F.linear(F.gelu(F.linear(ctx, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"]), approximate="tanh"),
                       sd["text_embedding.2.weight"], sd["text_embedding.2.bias"])
        return x, e, e0, ctx, grid

    def head(self, x, e):
        """Head.forward, model.py:388-401"""
        sd = self.sd
        m = (sd["head.modulation"].reshape(2, -1) + e).unbind(0)
        return F.linear(layer_norm(x, None, None, self.eps) * (1 + m[1]) + m[0], sd["head.head.weight"], sd["head.head.bias"])

    def unpatchify(self, y, grid):
        """model.py:633-656"""
        c = self.out_dim
        u = y.view(*grid, *self.patch_size, c)
        u = torch.einsum("fhwpqrc->cfphqwr", u)
        return u.reshape(c, *[i * j for i, j in zip(grid, self.patch_size)])

    def forward(self, latent, t, context):
        x, e, e0, ctx, grid = self.embed(latent, t, context)
        for blk in self.blocks:
            x = blk.forward(x, e0, grid, ctx)
        return self.unpatchify(self.head(x, e), grid)
