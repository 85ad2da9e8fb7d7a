"""QuantWanModel (wan_b200.quant_wanx; reference wan/quant_wanx.py:28-228) through the reference's own call sequence
(quant_generate.py:355-420): from_pretrained -> quant_layer_refactor -> load_quant_param_dict -> quantize_and_save_weight ->
hardware_forward_refactor -> set_init_done -> forward.

CPU: host logic on the oracle-backed fake backend, the FP tree against the full-DiT oracle.
GPU: the whole sequence on libb200q; the hardware forward against the fake-quant oracle."""
import json
import os

import pytest
import torch
from omegaconf import OmegaConf

import fake_backend

REGEX = r"text_embedding|time_embedding|time_projection|head\.head"
DIMS = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=64, text_len=32)


def _cfg():
    return OmegaConf.create({"remain_fp_regex": REGEX, "weight": {"n_bits": 8, "sym": False}, "act": {"n_bits": 8, "sym": True}})


def _ckpt_dir(tmp_path, seed=0):
    """a checkpoint directory in the released layout: config.json + a weight file"""
    from wan_b200.quant_wanx import QuantWanModel
    torch.manual_seed(seed)
    m = QuantWanModel(**DIMS)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(".bias"):
                p.normal_(0, 0.02, generator=g)
        m.head.head.weight.normal_(0, 0.05, generator=g)        # init zeroes it: the step output would be constant
    d = os.path.join(tmp_path, "ckpt")
    os.makedirs(d, exist_ok=True)
    json.dump({"_class_name": "WanModel", **DIMS, "model_type": "t2v"}, open(os.path.join(d, "config.json"), "w"))
    torch.save(m.state_dict(), os.path.join(d, "diffusion_pytorch_model.pt"))
    return d, {k: v.detach().clone() for k, v in m.state_dict().items()}


def _inputs(seed=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(16, 2, 8, 12, generator=g), torch.tensor([431.0]), torch.randn(20, 64, generator=g)


def test_fp_tree_matches_dit_oracle_cpu(tmp_path):
    from oracle import fakequant_oracle as O
    from wan_b200.quant_wanx import QuantWanModel
    d, sd = _ckpt_dir(str(tmp_path))
    m = QuantWanModel.from_pretrained(d).eval()
    assert (m.dim, m.ffn_dim, m.num_layers, m.text_len) == (256, 512, 2, 32)
    lat, t, ctx = _inputs()
    with torch.no_grad():
        y = m([lat], t, [ctx], seq_len=48)[0]
    ref = O.WanDiTOracle(sd, 256, 512, 2, 2, freq_dim=64, text_len=32, quant=False).forward(lat, t, ctx)
    assert y.shape == lat.shape and torch.allclose(y, ref, rtol=1e-4, atol=1e-5)


def test_reference_call_sequence_cpu(monkeypatch, tmp_path):
    """quant_layer_refactor / save+load_quant_param_dict / set_init_done / fake-quant forward on the fake backend"""
    fake_backend.install(monkeypatch)
    from oracle import fakequant_oracle as O
    from qdiff.base.quant_layer import QuantizedLinear
    from wan_b200.quant_wanx import QuantWanModel
    d, sd = _ckpt_dir(str(tmp_path))
    m = QuantWanModel.from_pretrained(d, quant_config=_cfg()).eval()
    m.quant_layer_refactor()
    assert isinstance(m.blocks[1].ffn[2], QuantizedLinear) and not isinstance(m.head.head, QuantizedLinear)
    m.save_quant_param_dict()
    params = {k: dict(v) for k, v in m.quant_param_dict.items()}
    assert len(params) == 2 * 10 * 2                               # w_ and a_quantizer of ten linears in two blocks
    m.load_quant_param_dict(params)
    m.set_init_done()
    lat, t, ctx = _inputs()
    with torch.no_grad():
        y = m([lat], t, [ctx], seq_len=48)[0]
    ref = O.WanDiTOracle(sd, 256, 512, 2, 2, freq_dim=64, text_len=32).forward(lat, t, ctx)
    cos = float((y.double().flatten() @ ref.double().flatten()) / (y.double().norm() * ref.double().norm()))
    assert cos >= 0.99999, cos


@pytest.mark.gpu
def test_quant_generate_sequence_gpu(dev, tmp_path):
    from oracle import fakequant_oracle as O
    from wan_b200.quant_wanx import QuantWanModel
    d, sd = _ckpt_dir(str(tmp_path))
    m = QuantWanModel.from_pretrained(d, quant_config=_cfg())
    m.to(dev)
    m.quant_layer_refactor()
    m.eval()
    m.save_quant_param_dict()                                       # stands in for the quant_params.pth of ptq_wanx.py
    ckpt = os.path.join(str(tmp_path), "quant_params.pth")
    torch.save({"_fsdp_wrapped_module." + k if i % 2 else k: v for i, (k, v) in enumerate(m.quant_param_dict.items())}, ckpt)
    loaded = torch.load(ckpt, map_location="cuda")
    clean = {k.replace("_fsdp_wrapped_module.", ""): v for k, v in loaded.items()}          # quant_generate.py:383-387
    m.load_quant_param_dict(clean)
    save_path = os.path.join(str(tmp_path), "int_weight.pt")
    m.quantize_and_save_weight(save_path=save_path)
    lat, t, ctx = _inputs()
    with torch.no_grad():
        y_sim = m([lat.to(dev)], t.to(dev), [ctx.to(dev)], seq_len=48)[0].cpu()        # algorithm-simulation path (:413-415)
    m.hardware_forward_refactor(load_path=save_path, seq_len=48)
    m.set_init_done()
    with torch.no_grad():
        y_hw = m([lat.to(dev)], t.to(dev), [ctx.to(dev)], 48)[0].cpu()
    ref = O.WanDiTOracle(sd, 256, 512, 2, 2, freq_dim=64, text_len=32).forward(lat, t, ctx)
    cos = lambda a, b: float((a.double().flatten() @ b.double().flatten()) / (a.double().norm() * b.double().norm()))
    assert cos(y_sim, ref) >= 0.9999 and cos(y_hw, ref) >= 0.999, (cos(y_sim, ref), cos(y_hw, ref))
