"""CPU suite: host logic of the qdiff mirror (module surgery, regex selection, parameter dicts, mixed precision),
with the kernel ops replaced by an oracle-backed fake (tests/fake_backend.py).  Mirrors how the reference's callers
use the surface: examples/Wan2.1/wan/quant_wanx.py:85-133, ptq_wanx.py:315-378, quant_generate.py:355-420."""
import os

import pytest
import torch
import torch.nn as nn
from omegaconf import OmegaConf

import fake_backend
from oracle import fakequant_oracle as O

SHIPPED_REGEX = (r"text_embedding|time_embedding|time_projection|head\.head|blocks\.\d+\.self_attn\.(?!q$)(?!k$)(?!v$)[^.]+"
                 r"|blocks\.\d+\.o|blocks\.\d+\.ffn.*|cross_attn")   # examples/Wan2.1/quant_configs/config.yaml:9


class Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q, self.k, self.v, self.o = (nn.Linear(d, d) for _ in range(4))


class Block(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.self_attn, self.cross_attn = Attn(d), Attn(d)
        self.ffn = nn.Sequential(nn.Linear(d, f), nn.GELU(approximate="tanh"), nn.Linear(f, d))


class Tiny(nn.Module):
    def __init__(self, d=32, f=64, n=2):
        super().__init__()
        self.text_embedding = nn.Sequential(nn.Linear(16, d), nn.GELU(), nn.Linear(d, d))
        self.blocks = nn.ModuleList(Block(d, f) for _ in range(n))
        self.head = nn.Module()
        self.head.head = nn.Linear(d, 8)
        self.quant_param_dict = {}


def _cfg(**extra):
    base = {"remain_fp_regex": SHIPPED_REGEX, "weight": {"n_bits": 8, "sym": False}, "act": {"n_bits": 8, "sym": True}}
    base.update(extra)
    return OmegaConf.create(base)


def _refactor(model, cfg):
    from qdiff.base.quant_model import quant_layer_refactor_
    from qdiff.utils import apply_func_to_submodules
    apply_func_to_submodules(model, class_type=nn.Linear, function=quant_layer_refactor_, name=None, parent_module=None,
                             quant_config=cfg, full_name=None, remain_fp_regex=cfg.remain_fp_regex)


def test_apply_func_to_submodules_names_and_return_dict():
    from qdiff.utils import apply_func_to_submodules
    m = Tiny()
    seen = apply_func_to_submodules(m, nn.Linear, lambda mod, full_name: full_name, return_d={}, full_name=None)
    assert "blocks.1.self_attn.q" in seen and seen["blocks.1.self_attn.q"] == "blocks.1.self_attn.q"
    assert "head.head" in seen and "blocks.0.ffn.2" in seen
    assert len(seen) == 2 + 2 * 10 + 1


def test_shipped_regex_quantizes_only_self_attn_qkv(monkeypatch):
    fake_backend.install(monkeypatch)
    from qdiff.base.quant_layer import QuantizedLinear
    m = Tiny()
    _refactor(m, _cfg())
    quantized = sorted(n for n, mod in m.named_modules() if isinstance(mod, QuantizedLinear))
    assert quantized == sorted(f"blocks.{i}.self_attn.{l}" for i in range(2) for l in "qkv")
    layer = m.blocks[0].self_attn.q
    assert layer.module_name == "blocks.0.self_attn.q" and layer.w_quantizer.module_name == layer.module_name
    assert layer.w_quantizer.init_done is True and layer.a_quantizer.init_done is False
    assert layer.w_quantizer.delta.shape == (32, 1) and layer.w_quantizer.zero_point.shape == (32, 1)
    # `weight` holds the fake-quantised view of the FP weight, exactly as the reference
    assert torch.equal(layer.weight.data, O.fake_quant_rows(layer.fp_weight.detach(), 8, False, dynamic=False))
    assert layer.fp_weight is layer.fp_module.weight and layer.bias is layer.fp_module.bias


def test_forward_matches_fake_quant_and_fp_bypass(monkeypatch):
    fake_backend.install(monkeypatch)
    m = Tiny()
    _refactor(m, _cfg(remain_fp_regex="text_embedding|head"))
    layer = m.blocks[1].ffn[0]
    x = torch.randn(2, 7, 32)
    y = layer(x)
    ref = O.quantized_linear_fake(x, layer.fp_weight.detach(), layer.bias.detach())
    assert y.shape == (2, 7, 64) and torch.allclose(y, ref, rtol=1e-5, atol=1e-5)
    layer.quant_mode = False
    assert torch.equal(layer(x), layer.fp_module(x))


def test_quant_param_dict_round_trip(monkeypatch, tmp_path):
    fake_backend.install(monkeypatch)
    from qdiff.base.base_quantizer import BaseQuantizer
    from qdiff.base.quant_model import load_quant_param_dict_, save_quant_param_dict_, set_init_done_
    from qdiff.utils import apply_func_to_submodules
    m = Tiny()
    _refactor(m, _cfg())
    apply_func_to_submodules(m, class_type=BaseQuantizer, function=set_init_done_)
    apply_func_to_submodules(m, class_type=BaseQuantizer, function=save_quant_param_dict_, full_name=None,
                             parent_module=None, model=m)
    keys = sorted(m.quant_param_dict)
    assert "blocks.0.self_attn.q.w_quantizer" in keys and "blocks.0.self_attn.q.a_quantizer" in keys
    assert set(m.quant_param_dict["blocks.0.self_attn.q.w_quantizer"]) == {"delta", "zero_point"}
    path = os.path.join(tmp_path, "quant_params.pth")
    torch.save(m.quant_param_dict, path)

    m2 = Tiny()
    m2.load_state_dict({k: v for k, v in Tiny().state_dict().items()}, strict=False)
    _refactor(m2, _cfg())
    loaded = torch.load(path)
    apply_func_to_submodules(m2, class_type=BaseQuantizer, function=load_quant_param_dict_, full_name=None,
                             parent_module=None, quant_param_dict=loaded, model=m2)
    a, b = m.blocks[1].self_attn.v.w_quantizer, m2.blocks[1].self_attn.v.w_quantizer
    assert torch.equal(a.delta, b.delta) and torch.equal(a.zero_point, b.zero_point)
    assert m2.quant_param_dict.keys() == m.quant_param_dict.keys()


def test_mixed_precision_bitwidth_refactor(monkeypatch):
    fake_backend.install(monkeypatch)
    from qdiff.base.quant_layer import QuantizedLinear
    from qdiff.base.quant_model import bitwidth_refactor_
    from qdiff.utils import apply_func_to_submodules
    cfg = OmegaConf.create({
        "remain_fp_regex": "text_embedding|head",
        "weight": {"n_bits": [4, 8], "sym": False, "i_bitwidth": 1},
        "act": {"n_bits": [4, 8], "sym": True, "i_bitwidth": 1},
        "mixed_precision": {"weight": {"layer_name_regex": ["cross_attn\\.o", "ffn", "self_attn|cross_attn"]},
                            "act": {"layer_name_regex": ["", "", ""]}},
    })
    m = Tiny()
    _refactor(m, cfg)
    apply_func_to_submodules(m, class_type=QuantizedLinear, function=bitwidth_refactor_, name=None, parent_module=None,
                             quant_config=cfg, full_name=None)
    ffn, sa, co = m.blocks[0].ffn[0], m.blocks[0].self_attn.k, m.blocks[0].cross_attn.o
    assert ffn.w_quantizer.n_bits == 4 and sa.w_quantizer.n_bits == 8
    assert co.quant_mode is False and ffn.quant_mode is True
    assert ffn.w_quantizer.delta_list.shape == (2, 64, 1)
    assert torch.equal(ffn.w_quantizer.delta, ffn.w_quantizer.delta_list[0])
    x = torch.randn(1, 5, 32)
    ref4 = O.quantized_linear_fake(x, ffn.fp_weight.detach(), ffn.bias.detach(), w_bits=4)
    assert torch.allclose(ffn(x), ref4, rtol=1e-5, atol=1e-5)
    assert torch.equal(ffn.weight.data, O.fake_quant_rows(ffn.fp_weight.detach(), 4, False, dynamic=False))


def test_method_sections_select_the_variant_layers(monkeypatch):
    """The shipped YAML's `viditq:` section with `layer_name_regex: ""` selects ViDiTQuantizedLinear for every quantized
    layer (quant_configs/config.yaml:19-21, quant_model.py:45-53); a layer without its PTQ state refuses to run."""
    fake_backend.install(monkeypatch)
    from qdiff.viditq.viditq_quant_layer import ViDiTQuantizedLinear
    m = Tiny()
    _refactor(m, _cfg(viditq={"alpha": 0.5665, "layer_name_regex": ""}))
    assert isinstance(m.blocks[0].self_attn.q, ViDiTQuantizedLinear)
    assert isinstance(m.blocks[0].self_attn.o, nn.Linear) and not isinstance(m.blocks[0].self_attn.o, ViDiTQuantizedLinear)
    with pytest.raises(RuntimeError):
        m.blocks[0].self_attn.q(torch.randn(1, 4, 32))      # channel_mask / rotation not set yet


def test_quantizers_refuse_to_run_without_a_gpu():
    import b200q
    from qdiff.base.base_quantizer import DynamicQuantizer
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    qz = DynamicQuantizer(OmegaConf.create({"n_bits": 8, "sym": True}))
    with pytest.raises(b200q.B200QError):
        qz(torch.randn(4, 16))


def test_attn_map_quantizer_row_group(monkeypatch):
    fake_backend.install(monkeypatch)
    from qdiff.base.quant_attn import QuantizedAttentionMapOpenSORA
    cfg = OmegaConf.create({"attn": {"qk": {"n_bits": 8, "sym": True, "reorder_file_path": None},
                                     "attn_map": {"n_bits": 8, "sym": False, "group": "row"}}})
    pm = QuantizedAttentionMapOpenSORA(cfg)
    rec = torch.load(os.path.join(os.path.dirname(__file__), "golden", "quant_attention.pt"))
    assert torch.equal(pm(rec["attn"].clone()), rec["attn_quant"])
