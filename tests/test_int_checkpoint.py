"""Integer-weight checkpoint (quant_wanx.py:137-185,221-228, quant_wanx_cuda.py:39-55): export from a
PTQ'd qdiff model, key schema, wrapper-prefix stripping (CPU, fake backend); load into the integer runtime and run (GPU)."""
import math
import os

import pytest
import torch
import torch.nn as nn
from omegaconf import OmegaConf

import fake_backend

REGEX = r"text_embedding|time_embedding|time_projection|head\.head|patch_embedding"


class _Norm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q, self.k, self.v, self.o = (nn.Linear(d, d) for _ in range(4))
        self.norm_q, self.norm_k = _Norm(d), _Norm(d)


class _Block(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.self_attn, self.cross_attn = _Attn(d), _Attn(d)
        self.norm3 = nn.LayerNorm(d)
        self.ffn = nn.Sequential(nn.Linear(d, f), nn.GELU(approximate="tanh"), nn.Linear(f, d))
        self.modulation = nn.Parameter(torch.randn(1, 6, d) / d ** 0.5)


class _Head(nn.Module):
    def __init__(self, d, o):
        super().__init__()
        self.head = nn.Linear(d, o)
        self.modulation = nn.Parameter(torch.randn(1, 2, d) / d ** 0.5)


class TinyWan(nn.Module):
    """Module tree with WanModel's parameter names (wan/modules/model.py:480-540)."""

    def __init__(self, d=256, f=512, layers=1, text_dim=64, freq_dim=64):
        super().__init__()
        self.patch_embedding = nn.Conv3d(16, d, kernel_size=(1, 2, 2), stride=(1, 2, 2))
        self.text_embedding = nn.Sequential(nn.Linear(text_dim, d), nn.GELU(approximate="tanh"), nn.Linear(d, d))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, d), nn.SiLU(), nn.Linear(d, d))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(d, d * 6))
        self.blocks = nn.ModuleList(_Block(d, f) for _ in range(layers))
        self.head = _Head(d, 64)
        self.quant_param_dict = {}


def _quantize(model, w_bits=8):
    from qdiff.base.quant_model import quant_layer_refactor_
    from qdiff.utils import apply_func_to_submodules
    cfg = OmegaConf.create({"remain_fp_regex": REGEX, "weight": {"n_bits": w_bits, "sym": False},
                            "act": {"n_bits": 8, "sym": True}})
    apply_func_to_submodules(model, class_type=nn.Linear, function=quant_layer_refactor_, name=None, parent_module=None,
                             quant_config=cfg, full_name=None, remain_fp_regex=cfg.remain_fp_regex)
    return model


def test_export_schema_cpu(monkeypatch):
    fake_backend.install(monkeypatch)
    from wan_b200 import int_checkpoint as IC
    from oracle import fakequant_oracle as O
    torch.manual_seed(0)
    m = _quantize(TinyWan(d=32, f=64))
    sd = IC.export_int_state_dict(m, device=torch.device("cpu"))
    assert sd[IC.META_KEY]["format"] == "b200q-int-weight"
    for name in ("blocks.0.self_attn.q", "blocks.0.cross_attn.v", "blocks.0.ffn.0", "blocks.0.ffn.2"):
        assert sd[name + ".weight"].dtype == torch.int8
        q, d, z = O.quant_rows(dict(m.named_modules())[name].fp_module.weight.detach(), 8, False, dynamic=False)
        assert torch.equal(sd[name + ".weight"].float(), q.clamp(-128, 127))
        assert torch.equal(sd[name + ".scale_weight"], d.flatten()) and torch.equal(sd[name + ".zp_weight"], z.flatten())
        assert name + ".bias" in sd
    assert not any("fp_module" in k or "fp_weight" in k or "w_quantizer" in k for k in sd if k != IC.META_KEY)
    assert sd["text_embedding.0.weight"].dtype == torch.float32          # remain_fp layers untouched
    assert torch.equal(sd["blocks.0.norm1.weight"], torch.ones(32))       # quant_wanx.py:170-175
    assert IC.strip_wrapper_prefixes("_fsdp_wrapped_module.blocks.0._fsdp_wrapped_module.ffn.0.weight") == "blocks.0.ffn.0.weight"


@pytest.mark.gpu
def test_save_load_roundtrip_gpu(dev, tmp_path):
    from wan_b200 import int_checkpoint as IC
    from wan_b200 import model as M
    torch.manual_seed(0)
    d, f = 256, 512
    m = _quantize(TinyWan(d=d, f=f, layers=2).to(dev))
    path = os.path.join(tmp_path, "int_weight.pt")
    IC.save_int_checkpoint(m, path)
    cfg = M.WanConfig(dim=d, ffn_dim=f, num_heads=2, num_layers=2, text_dim=64, freq_dim=64)
    dit = IC.load_int_checkpoint(cfg, path)
    # the loaded integer weights are exactly the layer's own integer state
    ref = M.QWeight.from_quantized_linear(m.blocks[1].ffn[0])
    assert torch.equal(dit.blocks[1].w_f0.codes, ref.codes) and torch.equal(dit.blocks[1].w_f0.delta, ref.delta)
    assert torch.equal(dit.blocks[1].w_f0.zp, ref.zp)
    q = M.QWeight.from_quantized_linear(m.blocks[0].self_attn.k)
    assert torch.equal(dit.blocks[0].w_qkv.codes[d:2 * d], q.codes)
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randn(16, 2, 8, 12, device=dev, generator=g)
    y = dit.forward(lat, torch.tensor([300.0], device=dev), torch.randn(20, 64, device=dev, generator=g))
    assert y.shape == lat.shape and torch.isfinite(y).all()


class _TinyQuantWan(TinyWan):
    pass


def test_quant_wan_mixin_surface_cpu(monkeypatch):
    """QuantWanModel's method surface (quant_wanx.py:80-185) through the mixin: refactor, param dict round trip, export."""
    fake_backend.install(monkeypatch)
    from wan_b200.quant_wanx import QuantWanMixin
    from qdiff.base.quant_layer import QuantizedLinear

    class M(QuantWanMixin, TinyWan):
        pass
    torch.manual_seed(0)
    m = M(d=32, f=64)
    m.convert_quant(OmegaConf.create({"remain_fp_regex": REGEX, "weight": {"n_bits": 8, "sym": False},
                                      "act": {"n_bits": 8, "sym": True}}))
    assert isinstance(m.blocks[0].ffn[0], QuantizedLinear) and not isinstance(m.text_embedding[0], QuantizedLinear)
    m.save_quant_param_dict()
    assert "blocks.0.self_attn.q.w_quantizer" in m.quant_param_dict
    saved = {k: dict(v) for k, v in m.quant_param_dict.items()}
    m.load_quant_param_dict(saved)
    m.set_init_done()
    assert m.blocks[0].ffn[2].a_quantizer.init_done is True
    cfg = m.wan_config()
    assert (cfg.dim, cfg.ffn_dim, cfg.num_layers, cfg.text_dim, cfg.freq_dim) == (32, 64, 1, 64, 64)


@pytest.mark.gpu
def test_hardware_forward_refactor_gpu(dev, tmp_path):
    """quantize_and_save_weight -> hardware_forward_refactor (quant_wanx.py:137-228): forward() runs on the integer
    runtime with WanModel's call convention, eager and CUDA-graph replay agree bit for bit."""
    from wan_b200.quant_wanx import QuantWanMixin

    class M(QuantWanMixin, TinyWan):
        pass
    torch.manual_seed(0)
    m = M(d=256, f=512, layers=2).to(dev)
    m.convert_quant(OmegaConf.create({"remain_fp_regex": REGEX, "weight": {"n_bits": 8, "sym": False},
                                      "act": {"n_bits": 8, "sym": True}}))
    path = os.path.join(tmp_path, "int_weight.pt")
    m.quantize_and_save_weight(path)
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randn(16, 2, 8, 12, device=dev, generator=g)
    ctx = torch.randn(20, 64, device=dev, generator=g)
    t = torch.tensor([300.0], device=dev)
    m.hardware_forward_refactor(path, seq_len=48, use_graph=False)
    y_eager = m([lat], t, [ctx], 48)[0].clone()
    m.hardware_forward_refactor(path, seq_len=48, use_graph=True)
    y_graph = m([lat], t, [ctx], 48)[0].clone()
    y_graph2 = m([lat], t, [ctx], 48)[0].clone()          # second call replays the captured graph
    assert y_eager.shape == lat.shape and torch.isfinite(y_eager).all()
    assert torch.equal(y_eager, y_graph) and torch.equal(y_graph, y_graph2)


def test_prefix_stripping_is_per_component():
    from wan_b200 import int_checkpoint as IC
    assert IC.strip_wrapper_prefixes("blocks.0.ffn.0.fp_module.bias") == "blocks.0.ffn.0.fp_module.bias"
    assert IC.strip_wrapper_prefixes("module.blocks.0._orig_mod.ffn.0.weight") == "blocks.0.ffn.0.weight"
    assert IC.strip_wrapper_prefixes("_fsdp_wrapped_module.head.modulation") == "head.modulation"


def test_export_w4_fp_and_variant_entries_cpu(monkeypatch):
    """4-bit layers export packed codes + in_features; layers the config keeps FP export their FP weight; no junk
    `fp_bias` keys; ViDiT-Q layers export channel_mask + rotation_sign."""
    fake_backend.install(monkeypatch)
    from wan_b200 import int_checkpoint as IC
    from qdiff.base.quant_model import quant_layer_refactor_
    from qdiff.utils import apply_func_to_submodules
    torch.manual_seed(0)
    m = _quantize(TinyWan(d=32, f=64), w_bits=4)
    sd = IC.export_int_state_dict(m, device=torch.device("cpu"))
    assert "blocks.0.ffn.0.weight_packed" in sd and "blocks.0.ffn.0.weight" not in sd
    assert sd[IC.META_KEY]["weight_bits"]["blocks.0.ffn.0"] == 4 and sd[IC.META_KEY]["in_features"]["blocks.0.ffn.2"] == 64
    assert not any(k.endswith("fp_bias") or "fp_module" in k for k in sd if k != IC.META_KEY)
    # shipped YAML: only self_attn q/k/v quantized, ViDiT-Q for every quantized layer
    cfg = OmegaConf.create({"remain_fp_regex": SHIPPED_FP, "weight": {"n_bits": 8, "sym": False}, "act": {"n_bits": 8, "sym": True},
                            "viditq": {"alpha": 0.5665, "layer_name_regex": ""}})
    m = TinyWan(d=128, f=64)
    apply_func_to_submodules(m, class_type=nn.Linear, function=quant_layer_refactor_, name=None, parent_module=None,
                             quant_config=cfg, full_name=None, remain_fp_regex=cfg.remain_fp_regex)
    _ptq_viditq(m)
    sd = IC.export_int_state_dict(m, device=torch.device("cpu"))
    assert sd["blocks.0.self_attn.q.weight"].dtype == torch.int8 and sd["blocks.0.self_attn.o.weight"].dtype == torch.float32
    assert sd["blocks.0.self_attn.k.channel_mask"].shape == (128,) and sd["blocks.0.self_attn.k.rotation_sign"].dtype == torch.int8
    assert "blocks.0.self_attn.k.rotation_matrix" not in sd and "blocks.0.ffn.0.scale_weight" not in sd


SHIPPED_FP = (r"text_embedding|time_embedding|time_projection|head\.head|blocks\.\d+\.self_attn\.(?!q$)(?!k$)(?!v$)[^.]+"
              r"|blocks\.\d+\.o|blocks\.\d+\.ffn.*|cross_attn")


def _ptq_viditq(m, seed=3):
    """What ptq_wanx.py:330-345 does per ViDiT-Q layer: channel mask from the calibration abs-max, a fresh rotation,
    re-quantised scaled + rotated weights."""
    from qdiff.viditq.viditq_quant_layer import ViDiTQuantizedLinear
    g = torch.Generator().manual_seed(seed)
    n = 0
    for mod in m.modules():
        if isinstance(mod, ViDiTQuantizedLinear):
            dev = mod.fp_module.weight.device
            mod.get_channel_mask((torch.rand(mod.in_features, generator=g) * 3 + 0.2).clamp_min(1e-3).to(dev))
            mod.channel_mask = mod.channel_mask.detach()
            mod.get_rotation_matrix()
            mod.update_quantized_weight_rotated_and_scaled()
            n += 1
    return n


@pytest.mark.gpu
def test_w4_checkpoint_roundtrip_gpu(dev, tmp_path):
    from wan_b200 import int_checkpoint as IC
    from wan_b200 import model as M
    torch.manual_seed(0)
    d, f = 256, 512
    m = _quantize(TinyWan(d=d, f=f, layers=2).to(dev), w_bits=4)
    path = os.path.join(tmp_path, "int_weight_w4.pt")
    sd = IC.save_int_checkpoint(m, path)
    assert "blocks.1.ffn.0.weight_packed" in sd
    cfg = M.WanConfig(dim=d, ffn_dim=f, num_heads=2, num_layers=2, text_dim=64, freq_dim=64)
    dit = IC.load_int_checkpoint(cfg, path)
    ref = M.QWeight.from_quantized_linear(m.blocks[1].ffn[0])
    w = dit.blocks[1].w_f0
    assert w.codes is None and w.n_bits == 4 and w.K == d and torch.equal(w.packed, ref.packed) and torch.equal(w.delta, ref.delta)
    qkv = dit.blocks[0].w_qkv                                   # packed rows concatenate
    assert qkv is not None and torch.equal(qkv.packed[d:2 * d], M.QWeight.from_quantized_linear(m.blocks[0].self_attn.k).packed)
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randn(16, 2, 8, 12, device=dev, generator=g)
    ctx = torch.randn(20, 64, device=dev, generator=g)
    t = torch.tensor([300.0], device=dev)
    y = dit.forward(lat, t, ctx)
    # same step built directly from the layers' own integer state
    blocks = []
    for b in m.blocks:
        w = {n: M.QWeight.from_quantized_linear(dict(b.named_modules())[n]) for n in M.LINEARS}
        for k in ("self_attn.norm_q.weight", "self_attn.norm_k.weight", "cross_attn.norm_q.weight", "cross_attn.norm_k.weight",
                  "norm3.weight", "norm3.bias", "modulation"):
            w[k] = dict(b.named_parameters())[k].detach().float().contiguous()
        blocks.append(M.WanBlockQ(cfg, w))
    y2 = M.WanDiTQ(cfg, blocks, dit.fp).forward(lat, t, ctx)
    assert torch.isfinite(y).all() and torch.equal(y, y2)


@pytest.mark.gpu
def test_viditq_shipped_config_hardware_forward_vs_oracle_gpu(dev, tmp_path):
    """The shipped flow for the shipped YAML (quant_configs/config.yaml: ViDiT-Q on blocks.N.self_attn.{q,k,v}, everything
    else FP): PTQ -> quantize_and_save_weight -> hardware_forward_refactor -> forward, against the reference's forward
    formulas (viditq_quant_layer.py:52-73: x*mask, fp64 rotation, fake-quant activations, F.linear on the layer's
    fake-quant weight) evaluated with the oracle on the CPU.  A runtime that drops channel_mask / rotation (reference
    defect B-4, round-1 finding) lands far below the bound."""
    from oracle import fakequant_oracle as O
    from qdiff.viditq.viditq_quant_layer import ViDiTQuantizedLinear
    from wan_b200.quant_wanx import QuantWanMixin

    class Mx(QuantWanMixin, TinyWan):
        pass
    torch.manual_seed(0)
    d, f, layers = 256, 512, 2
    m = Mx(d=d, f=f, layers=layers).to(dev)
    with torch.no_grad():
        for b in m.blocks:                                      # an outlier input channel, the case ViDiT-Q exists for
            b.self_attn.q.weight[:, 5] *= 6
    sd_fp = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    m.convert_quant(OmegaConf.create({"remain_fp_regex": SHIPPED_FP, "weight": {"n_bits": 8, "sym": False},
                                      "act": {"n_bits": 8, "sym": True}, "viditq": {"alpha": 0.5665, "layer_name_regex": ""}}))
    assert _ptq_viditq(m) == 3 * layers
    layers_by_name = {n: mod for n, mod in m.named_modules() if isinstance(mod, ViDiTQuantizedLinear)}
    assert sorted(layers_by_name) == sorted(f"blocks.{i}.self_attn.{x}" for i in range(layers) for x in "qkv")
    path = os.path.join(tmp_path, "int_weight_viditq.pt")
    m.quantize_and_save_weight(path)

    def lin(name, x, default):
        w, b = sd_fp[name + ".weight"], sd_fp[name + ".bias"]
        if name in layers_by_name:
            L = layers_by_name[name]
            xr = (x.double() * L.channel_mask.double().cpu().reshape(1, -1)) @ L.rotation_matrix.double().cpu()
            return torch.nn.functional.linear(O.fake_quant_rows(xr.float(), 8, True, True), L.weight.data.float().cpu(), b)
        return torch.nn.functional.linear(x, w, b)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(16, 2, 8, 12, generator=g)
    ctx = torch.randn(20, 64, generator=g)
    t = torch.tensor([300.0])
    ref = O.WanDiTOracle(sd_fp, d, f, 2, layers, freq_dim=64, text_len=512, lin=lin).forward(lat, t, ctx)
    m.hardware_forward_refactor(path, seq_len=48, use_graph=False)
    blk = m._b200_dit.blocks[0]
    assert blk.w_qkv is None and blk.lin["self_attn.q"].pre is not None            # per-layer transforms: no shared codes
    out = m([lat.to(dev)], t.to(dev), [ctx.to(dev)], 48)[0].cpu()
    a, b = out.double().flatten(), ref.double().flatten()
    cos = float((a @ b) / (a.norm() * b.norm()))
    assert cos >= 0.999, cos
    # dropping the transform is NOT within tolerance (the check has teeth)
    for bq in m._b200_dit.blocks:
        for n in ("self_attn.q", "self_attn.k", "self_attn.v"):
            bq.lin[n].pre = None
    bad = m([lat.to(dev)], t.to(dev), [ctx.to(dev)], 48)[0].cpu().double().flatten()
    assert float((bad @ b) / (bad.norm() * b.norm())) < 0.995 < cos
