"""-m gpu: block / step parity at the BASELINE.json model dims, at a reduced token count the CPU
oracle finishes in seconds: Wan2.1-1.3B (D=1536, F=8960, H=12) and Wan2.1-14B (D=5120, F=13824, H=40), plain inputs and
the heavy-tailed variant (0.1 % of the channels x50), bf16 attention and the int8 attention kernel;
a 3-block DiT step against the chained oracle blocks; the shipped YAML's layer selection (only self_attn q/k/v quantized).

Tolerance (north_star): cosine >= 0.999 per block output; max relative error (of the output range) stated per case."""
import math

import pytest
import torch

import b200q
from oracle import fakequant_oracle as O
from wan_b200 import model as M

pytestmark = pytest.mark.gpu

DIMS = {"1.3B": (1536, 8960, 12), "14B": (5120, 13824, 40)}


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()))


def _inputs(dim, grid, T, outliers, seed=1):
    g = torch.Generator().manual_seed(seed)
    L = grid[0] * grid[1] * grid[2]
    x = torch.randn(L, dim, generator=g)
    if outliers:                                        # 0.1 % of the channels x50
        idx = torch.randperm(dim, generator=g)[:max(1, dim // 1000)]
        x[:, idx] *= 50
    e = torch.randn(6, dim, generator=g) * 0.1
    ctx = torch.randn(T, dim, generator=g)
    return x, e, ctx


@pytest.mark.parametrize("model", ["1.3B", "14B"])
@pytest.mark.parametrize("outliers", [False, True])
@pytest.mark.parametrize("attn", ["bf16", "int8"])
def test_block_parity_at_baseline_dims(dev, model, outliers, attn):
    dim, ffn, heads = DIMS[model]
    grid, T = (2, 6, 8), 40
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=heads, num_layers=1)
    p = O.make_block_params(dim, ffn, seed=0)
    x, e, ctx = _inputs(dim, grid, T, outliers)
    ref = O.WanBlockOracle(p, dim, ffn, heads, attn_quant=dict(mode="rowstep") if attn == "int8" else None
                           ).forward(x.clone(), e, grid, ctx)
    blk = M.WanBlockQ.from_fp_params(cfg, p, attn_quant=(attn == "int8"))
    cos, sin = M.rope_table(dim // heads, grid, dev)
    out = blk.forward(x.clone().to(dev), e.to(dev), ctx.to(dev), cos, sin).cpu()
    c = _cos(out, ref)
    # the residual stream dominates the block OUTPUT (x + ...): also compare what the block ADDED
    c_delta = _cos(out - x, ref - x)
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert c >= 0.999 and c_delta >= 0.999, (c, c_delta)
    assert rel <= 5e-2, rel


@pytest.mark.parametrize("model", ["1.3B", "14B"])
def test_block_weight_and_activation_codes_bit_exact_at_baseline_dims(dev, model):
    """first quantized linear of the block: LN+modulate activation codes / delta and all weight codes vs the oracle"""
    dim, ffn, heads = DIMS[model]
    p = O.make_block_params(dim, ffn, seed=2)
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=heads, num_layers=1)
    blk = M.WanBlockQ.from_fp_params(cfg, p)
    for name, w in (("ffn.0", blk.w_f0), ("ffn.2", blk.w_f2), ("cross_attn.q", blk.w_cq)):
        q, d, z = O.quant_rows(p[name + ".weight"], 8, False, dynamic=False)
        # the reference's float codes reach +128 on an exact double-rounding tie; int8 storage holds +127
        assert int((q > 127).sum()) <= 8
        assert torch.equal(w.codes.cpu().float(), q.clamp(-128, 127)) and torch.equal(w.delta.cpu(), d.flatten()) and torch.equal(w.zp.cpu(), z.flatten())
    x, e, _ = _inputs(dim, (2, 4, 4), 8, True)
    em = (p["modulation"].reshape(6, -1) + e)
    h = O.layer_norm(x, None, None, 1e-6) * (1 + em[1]) + em[0]
    qa, da, rs, y = b200q.ln_mod_quant(x.to(dev), 1e-6, shift=em[0].to(dev), scale=em[1].to(dev), y_dtype=torch.float32)
    assert torch.allclose(y.cpu(), h, rtol=2e-5, atol=2e-5)
    qo, do, _ = O.quant_rows(y.cpu(), 8, True, True)               # codes of the kernel's own normalised activations
    assert torch.equal(qa.cpu().float(), qo) and torch.equal(da.cpu(), do.flatten())


def _wan_state_dict(cfg, seed=0):
    """FP state dict with WanModel's parameter names and init statistics (model.py:658-680; biases re-drawn)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(cfg.num_layers):
        for k, v in O.make_block_params(cfg.dim, cfg.ffn_dim, seed=seed * 100 + i).items():
            sd[f"blocks.{i}.{k}"] = v

    def lin(name, o, i, std=None):
        a = math.sqrt(6.0 / (i + o))
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * a if std is None else torch.randn(o, i, generator=g) * std
        sd[name + ".bias"] = torch.randn(o, generator=g) * 0.02
    pdim = cfg.in_dim * math.prod(cfg.patch_size)
    lin("patch_embedding", cfg.dim, pdim)
    sd["patch_embedding.weight"] = sd["patch_embedding.weight"].reshape(cfg.dim, cfg.in_dim, *cfg.patch_size)
    lin("text_embedding.0", cfg.dim, cfg.text_dim, 0.02); lin("text_embedding.2", cfg.dim, cfg.dim, 0.02)
    lin("time_embedding.0", cfg.dim, cfg.freq_dim, 0.02); lin("time_embedding.2", cfg.dim, cfg.dim, 0.02)
    lin("time_projection.1", 6 * cfg.dim, cfg.dim)
    lin("head.head", cfg.out_dim * math.prod(cfg.patch_size), cfg.dim, 0.02)
    sd["head.modulation"] = torch.randn(1, 2, cfg.dim, generator=g) / cfg.dim ** 0.5
    return sd


@pytest.mark.parametrize("dim,ffn,heads", [(1536, 8960, 12), (256, 512, 2)])
def test_three_block_dit_step_vs_chained_oracle(dev, dim, ffn, heads):
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=heads, num_layers=3, text_dim=96, text_len=24, freq_dim=64)
    sd = _wan_state_dict(cfg)
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(16, 2, 12, 16, generator=g)
    ctx = torch.randn(17, 96, generator=g)
    t = torch.tensor([431.0])
    ref = O.WanDiTOracle(sd, dim, ffn, heads, 3, freq_dim=64, text_len=24).forward(lat, t, ctx)
    dit = M.WanDiTQ.from_fp_state_dict(cfg, sd)
    out = dit.forward(lat.to(dev), t.to(dev), ctx.to(dev)).cpu()
    c = _cos(out, ref)
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert out.shape == lat.shape and c >= 0.999, c
    assert rel <= 8e-2, rel             # bf16 embeddings / head (the reference autocasts them too) on top of the block tolerance
    # CUDA-graph replay and the CFG-batched step (cond/uncond stacked along the rows) give the same numbers
    graphed = M.GraphedDiT(dit)
    assert torch.equal(graphed(lat.to(dev), t.to(dev), ctx.to(dev)).cpu(), out) and graphed.failed is None
    ctx2 = torch.randn(9, 96, generator=g)
    both = dit.forward(lat.to(dev), t.to(dev), [ctx.to(dev), ctx2.to(dev)]).cpu()
    assert both.shape == (2,) + tuple(lat.shape)
    assert torch.equal(both[0], out)
    assert torch.equal(both[1], dit.forward(lat.to(dev), t.to(dev), ctx2.to(dev)).cpu())


SHIPPED_FP = r"text_embedding|time_embedding|time_projection|head\.head|blocks\.\d+\.self_attn\.(?!q$)(?!k$)(?!v$)[^.]+|blocks\.\d+\.o|blocks\.\d+\.ffn.*|cross_attn"


def test_shipped_yaml_layer_selection_vs_oracle(dev):
    """quant_configs/config.yaml:9: only blocks.N.self_attn.{q,k,v} are quantized, o / cross_attn / ffn stay FP."""
    import re
    dim, ffn, heads = 1536, 8960, 12
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=heads, num_layers=2, text_dim=96, text_len=24, freq_dim=64)
    sd = _wan_state_dict(cfg, seed=1)
    fp_re = re.compile(SHIPPED_FP)

    def lin(name, x, default):
        if fp_re.search(name):
            layer = name.split(".", 2)[2]
            pre = name[:len(name) - len(layer)]
            return torch.nn.functional.linear(x, sd[pre + layer + ".weight"], sd[pre + layer + ".bias"])
        return default()
    g = torch.Generator().manual_seed(6)
    lat = torch.randn(16, 2, 8, 12, generator=g)
    ctx = torch.randn(11, 96, generator=g)
    t = torch.tensor([77.0])
    ref = O.WanDiTOracle(sd, dim, ffn, heads, 2, freq_dim=64, text_len=24, lin=lin).forward(lat, t, ctx)
    dit = M.WanDiTQ.from_fp_state_dict(cfg, sd, remain_fp_regex=SHIPPED_FP)
    blk = dit.blocks[0]
    assert isinstance(blk.w_o, M.FPWeight) and isinstance(blk.w_f0, M.FPWeight) and isinstance(blk.w_cq, M.FPWeight)
    assert isinstance(blk.w_qkv, M.QWeight) and blk.w_ckv is None and isinstance(blk.lin["cross_attn.k"], M.FPWeight)
    out = dit.forward(lat.to(dev), t.to(dev), ctx.to(dev)).cpu()
    c = _cos(out, ref)
    assert c >= 0.999, c
