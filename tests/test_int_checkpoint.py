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
