"""SmoothQuant / QuaRot / ViDiT-Q layer variants against tests/golden/variant_layers.pt, the outputs
of the imported reference layers (oracle/gen_golden_layers.py).

CPU part: host logic + arithmetic through the oracle-backed fake backend — fake-quant weight, delta, zero_point BIT-EXACT;
layer output within 2e-4 (integer GEMM algebra vs the reference's fp32 F.linear on dequantised operands).
GPU part (-m gpu): the same layers on libb200q."""
import os

import pytest
import torch
import torch.nn as nn
from omegaconf import OmegaConf

import fake_backend

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "variant_layers.pt")
CASES = ["smooth_quant_w8", "smooth_quant_w4", "quarot_w8", "quarot_w4", "viditq_w8", "viditq_w4"]


def _build(rec, name, device="cpu"):
    from qdiff.smooth_quant.sq_quant_layer import SQQuantizedLinear
    from qdiff.quarot.quarot_quant_layer import QuarotQuantizedLinear
    from qdiff.viditq.viditq_quant_layer import ViDiTQuantizedLinear
    cls = {"smooth_quant": SQQuantizedLinear, "quarot": QuarotQuantizedLinear, "viditq": ViDiTQuantizedLinear}[name.rsplit("_", 1)[0]]
    cout, cin = rec["weight"].shape
    fp = nn.Linear(cin, cout)
    with torch.no_grad():
        fp.weight.copy_(rec["weight"]); fp.bias.copy_(rec["bias"])
    fp = fp.to(device)
    layer = cls(cin, cout, True, None, OmegaConf.create(rec["cfg"]), fp)
    if hasattr(layer, "channel_mask"):
        layer.get_channel_mask(rec["act_mask"].to(device))
        if str(device) == "cpu":
            assert torch.equal(layer.channel_mask, rec["channel_mask"])
        else:                               # pow() on the device differs from the host libm by an ulp
            assert torch.allclose(layer.channel_mask.cpu(), rec["channel_mask"], rtol=1e-5)
            layer.channel_mask = rec["channel_mask"].to(device)
    if hasattr(layer, "rotation_matrix"):
        layer.rotation_matrix = rec["rotation_matrix"].to(device)
    if name.startswith("smooth_quant"):
        layer.update_quantized_weight_scaled()
    elif name.startswith("quarot"):
        layer.update_quantized_weight_rotated()
    else:
        layer.update_quantized_weight_rotated_and_scaled()
    return layer


def _check(layer, rec, device="cpu"):
    assert torch.equal(layer.w_quantizer.delta.cpu(), rec["w_delta"])
    assert torch.equal(layer.w_quantizer.zero_point.cpu(), rec["w_zero_point"])
    # int8 storage saturates the asymmetric code +128 (reachable only on an exact double rounding tie) to
    # +127: at most a couple of weights may sit one step below the reference's float code
    dw = (layer.weight.data.cpu() - rec["fq_weight"]).abs()
    assert int((dw > 0).sum()) <= 2 and bool((dw <= rec["w_delta"] * 1.0001).all())
    with torch.no_grad():
        y = layer(rec["x"].to(device)).float().cpu()
    ref = rec["y"]
    assert y.shape == ref.shape
    # activation codes can flip by one step where the fp32 Hadamard transform and the reference's fp64 matmul round
    # differently; one flipped code moves an output by delta_a*|w| ~ 1e-3 of the output scale
    rel = float((y - ref).abs().max() / ref.abs().max())
    # 4-bit: a weight whose reference code is the unrepresentable +8 (exact double tie) is stored as +7, one 4-bit
    # step (1/15 of the row range) off - measured 6.3e-3 on viditq_w4
    assert rel <= (5e-3 if rec["w_bits"] == 8 else 2e-2), rel
    cos = float((y.double().flatten() @ ref.double().flatten()) / (y.double().norm() * ref.double().norm()))
    assert cos >= 0.99999, cos


@pytest.mark.parametrize("name", CASES)
def test_variant_layer_cpu(monkeypatch, name):
    fake_backend.install(monkeypatch)
    rec = torch.load(GOLDEN)[name]
    layer = _build(rec, name)
    _check(layer, rec)
    # quant_mode False -> the FP module, untouched (sq_quant_layer.py:50-51)
    layer.quant_mode = False
    assert torch.equal(layer(rec["x"]), torch.nn.functional.linear(rec["x"], rec["weight"], rec["bias"]))


def test_hadamard_constructions():
    from qdiff.quarot import quarot_utils as U
    for n, K in ((1536, 12), (8960, 140), (5120, 20), (13824, 108), (4096, 1), (1152, 36)):
        assert U.hadamard_factor(n)[0] == K
    for K in (12, 20, 28, 36, 40, 108, 140):
        H = U.base_hadamard(K)
        assert torch.equal(H.abs(), torch.ones_like(H)) and torch.equal(H @ H.t(), K * torch.eye(K, dtype=torch.float64))
    R = U.random_hadamard_matrix(96, "cpu")
    assert torch.allclose(R @ R.t(), torch.eye(96, dtype=torch.float64), atol=1e-12)
    x = torch.randn(7, 96, dtype=torch.float64)
    assert torch.allclose(U.matmul_hadU(x), x @ U.matmul_hadU(torch.eye(96, dtype=torch.float64)), atol=1e-12)
    assert torch.allclose(U.matmul_hadUt(U.matmul_hadU(x)), x, atol=1e-12)       # orthogonal: transpose inverts


def test_load_quant_param_dict_rebuilds_variant_weights(monkeypatch):
    """quant_model.py:138-160: loading re-derives the rotated/scaled weights from channel_mask (+ a fresh rotation)."""
    fake_backend.install(monkeypatch)
    from qdiff.base.quant_model import load_quant_param_dict_, save_quant_param_dict_
    rec = torch.load(GOLDEN)["viditq_w8"]
    layer = _build(rec, "viditq_w8")

    class Holder:
        quant_param_dict = {}
    h = Holder()
    save_quant_param_dict_(layer.w_quantizer, "l.w_quantizer", layer, h)
    entry = h.quant_param_dict["l.w_quantizer"]
    assert entry["rotation_matrix"] is None and torch.equal(entry["channel_mask"], rec["channel_mask"])
    fresh = _build(rec, "viditq_w8")
    fresh.channel_mask = None
    load_quant_param_dict_(fresh.w_quantizer, "l.w_quantizer", fresh, h.quant_param_dict, h)
    assert torch.equal(fresh.channel_mask, rec["channel_mask"]) and fresh.rotation_matrix is not None
    with torch.no_grad():
        y = fresh(rec["x"])
    cos = float((y.double().flatten() @ rec["y"].double().flatten()) / (y.double().norm() * rec["y"].double().norm()))
    assert cos >= 0.999          # a different random rotation, the same function


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_variant_layer_gpu(dev, name):
    rec = torch.load(GOLDEN)[name]
    layer = _build(rec, name, dev)
    _check(layer, rec, dev)
