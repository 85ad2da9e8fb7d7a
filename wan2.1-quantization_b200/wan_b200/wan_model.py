"""WanModelFP — a self-contained torch module tree with the parameter names, constructor arguments and call convention of
the reference's `WanModel` (ViDiT-Q/examples/Wan2.1/wan/modules/model.py:424-680), t2v variant.

It is the HOST of the quantization plugin surface when the reference's own `wan` package is not importable (it needs
easydict / xformers / diffusers / flash-attn at import time): `quant_layer_refactor` swaps its `nn.Linear`s for the
`qdiff` mirror layers, calibration hooks attach to it, `quantize_and_save_weight` exports it and
`hardware_forward_refactor` replaces its forward by the integer runtime (wan_b200.model.WanDiTQ).  Its own forward is
the "algorithm simulation" path of quant_generate.py (`if_hardware = False`, :413-415): plain torch glue around whatever
linear layers the tree currently holds.  Samples are processed one at a time as [L, D] matrices (the reference pads a
batch to `seq_len` and loops over samples inside rope_apply, model.py:51-70); semantics follow the Ulysses variant of
self-attention (xdit_context_parallel.py:155-192), i.e. the in-file version with its q path repaired (SURVEY B-1)."""
from __future__ import annotations

import json
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .model import WanConfig, rope_table, sinusoidal_embedding_1d


class WanRMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + self.eps)).type_as(x) * self.weight


class WanLayerNorm(nn.LayerNorm):
    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)

    def forward(self, x):
        return super().forward(x.float()).type_as(x)


def _rotate(x, cos, sin):
    """3-axis RoPE on adjacent channel pairs of every head: x [L, n, d]; cos/sin [L, d/2] (model.py:43-70)."""
    L, n, d = x.shape
    xr = x.float().reshape(L, n, d // 2, 2)
    re, im = xr[..., 0], xr[..., 1]
    c, s = cos.view(L, 1, -1), sin.view(L, 1, -1)
    return torch.stack((re * c - im * s, re * s + im * c), dim=-1).reshape(L, n, d).type_as(x)


def _attend(q, k, v):
    """q [Lq, n, d], k, v [Lk, n, d] -> [Lq, n*d] (attention.py:171-178)"""
    o = F.scaled_dot_product_attention(q.transpose(0, 1).unsqueeze(0), k.transpose(0, 1).unsqueeze(0), v.transpose(0, 1).unsqueeze(0))
    return o.squeeze(0).transpose(0, 1).flatten(1)


class WanSelfAttention(nn.Module):
    def __init__(self, dim, num_heads, qk_norm=True, eps=1e-6):
        super().__init__()
        assert dim % num_heads == 0
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def forward(self, x, cos, sin):
        L, n, d = x.shape[0], self.num_heads, self.head_dim
        q = _rotate(self.norm_q(self.q(x)).view(L, n, d), cos, sin)
        k = _rotate(self.norm_k(self.k(x)).view(L, n, d), cos, sin)
        return self.o(_attend(q, k, self.v(x).view(L, n, d)))


class WanT2VCrossAttention(WanSelfAttention):
    def forward(self, x, context):
        n, d = self.num_heads, self.head_dim
        q = self.norm_q(self.q(x)).view(-1, n, d)
        k = self.norm_k(self.k(context)).view(-1, n, d)
        return self.o(_attend(q, k, self.v(context).view(-1, n, d)))


class WanAttentionBlock(nn.Module):
    def __init__(self, dim, ffn_dim, num_heads, qk_norm=True, cross_attn_norm=False, eps=1e-6):
        super().__init__()
        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, qk_norm, eps)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True) if cross_attn_norm else nn.Identity()
        self.cross_attn = WanT2VCrossAttention(dim, num_heads, qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    def forward(self, x, e0, context, cos, sin):
        """x [L, D] fp32 residual stream, e0 [6, D] fp32 (model.py:293-370)"""
        e = (self.modulation.reshape(6, -1).float() + e0).unbind(0)
        x = x + self.self_attn(self.norm1(x).float() * (1 + e[1]) + e[0], cos, sin).float() * e[2]
        x = x + self.cross_attn(self.norm3(x), context).float()
        x = x + self.ffn(self.norm2(x).float() * (1 + e[4]) + e[3]).float() * e[5]
        return x


class Head(nn.Module):
    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, math.prod(patch_size) * out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)

    def forward(self, x, e):
        m = (self.modulation.reshape(2, -1).float() + e.float()).unbind(0)
        return self.head(self.norm(x).float() * (1 + m[1]) + m[0])


class WanModelFP(nn.Module):
    """Constructor arguments of WanModel.__init__ (model.py:434-452); `model_type` must be 't2v'."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192, freq_dim=256,
                 text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1), qk_norm=True,
                 cross_attn_norm=True, eps=1e-6):
        super().__init__()
        if model_type != "t2v":
            raise NotImplementedError("WanModelFP covers the text-to-video path (the quantized hot path of SURVEY §8)")
        self.model_type, self.patch_size, self.text_len, self.in_dim, self.dim = model_type, tuple(patch_size), text_len, in_dim, dim
        self.ffn_dim, self.freq_dim, self.text_dim, self.out_dim, self.num_heads = ffn_dim, freq_dim, text_dim, out_dim, num_heads
        self.num_layers, self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = num_layers, window_size, qk_norm, cross_attn_norm, eps
        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, dim), nn.GELU(approximate="tanh"), nn.Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        self.blocks = nn.ModuleList(WanAttentionBlock(dim, ffn_dim, num_heads, qk_norm, cross_attn_norm, eps) for _ in range(num_layers))
        self.head = Head(dim, out_dim, self.patch_size, eps)
        self.init_weights()

    def init_weights(self):
        """model.py:658-680: xavier-uniform linears with zero bias, N(0, 0.02) text / time embeddings, zero head."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for m in list(self.text_embedding.modules()) + list(self.time_embedding.modules()):
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.02)
        nn.init.zeros_(self.head.head.weight)

    def config(self) -> WanConfig:
        return WanConfig(dim=self.dim, ffn_dim=self.ffn_dim, num_heads=self.num_heads, num_layers=self.num_layers,
                         in_dim=self.in_dim, out_dim=self.out_dim, text_dim=self.text_dim, text_len=self.text_len,
                         freq_dim=self.freq_dim, patch_size=self.patch_size, eps=self.eps)

    def _forward_one(self, u, t, ctx):
        dev = self.patch_embedding.weight.device
        x = self.patch_embedding(u.unsqueeze(0).to(self.patch_embedding.weight.dtype))
        grid = tuple(x.shape[2:])
        x = x.flatten(2).transpose(1, 2)[0].float()                                        # [L, D]
        e = self.time_embedding(sinusoidal_embedding_1d(self.freq_dim, t.reshape(-1)[:1]).float().to(dev))    # [1, D]
        e0 = self.time_projection(e).view(6, self.dim).float()
        c = torch.zeros(self.text_len, self.text_dim, device=dev, dtype=ctx.dtype)
        c[:ctx.shape[0]] = ctx
        c = self.text_embedding(c)
        cos, sin = rope_table(self.dim // self.num_heads, grid, dev)
        for blk in self.blocks:
            x = blk(x, e0, c, cos, sin)
        y = self.head(x, e)
        p = self.patch_size
        y = torch.einsum("fhwpqrc->cfphqwr", y.view(*grid, *p, self.out_dim))
        return y.reshape(self.out_dim, *[i * j for i, j in zip(grid, p)]).float()

    def forward(self, x, t, context, seq_len=None, clip_fea=None, y=None):
        """x: list of [C, F, H, W]; t [B]; context: list of [T, text_dim] -> list of [C, F, H, W] fp32 (model.py:539-631)"""
        if clip_fea is not None or y is not None:
            raise NotImplementedError("image conditioning (i2v) is outside the quantized t2v path")
        return [self._forward_one(u, t[i:i + 1], c) for i, (u, c) in enumerate(zip(x, context))]

    @classmethod
    def from_pretrained(cls, ckpt_dir, **extra):
        """`ckpt_dir` with the layout of the released Wan2.1 checkpoints: config.json (constructor arguments) +
        diffusion_pytorch_model*.safetensors (or a .pt / .pth state dict).  `extra` (e.g. quant_config) goes to the
        constructor, as QuantWanModel.from_pretrained(ckpt_dir, quant_config=...) does (quant_generate.py:358-360)."""
        cfg = {}
        cfg_path = os.path.join(ckpt_dir, "config.json")
        if os.path.exists(cfg_path):
            raw = json.load(open(cfg_path))
            known = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
                     "num_heads", "num_layers", "window_size", "qk_norm", "cross_attn_norm", "eps")
            cfg = {k: raw[k] for k in known if k in raw}
        model = cls(**cfg, **extra)
        sd = {}
        files = sorted(f for f in os.listdir(ckpt_dir) if f.endswith((".safetensors", ".pt", ".pth", ".bin")))
        if not files:
            raise FileNotFoundError(f"no weight file (*.safetensors | *.pt | *.pth) in {ckpt_dir}")
        for f in files:
            path = os.path.join(ckpt_dir, f)
            if f.endswith(".safetensors"):
                from safetensors.torch import load_file
                sd.update(load_file(path))
            else:
                part = torch.load(path, map_location="cpu", weights_only=True)
                sd.update(part.get("state_dict", part) if isinstance(part, dict) else part)
        missing, unexpected = model.load_state_dict(sd, strict=False)
        if missing:
            raise RuntimeError(f"checkpoint in {ckpt_dir} lacks {len(missing)} parameters, e.g. {missing[:3]}")
        return model
