"""Smooth scale + randomized Hadamard rotation fused into the per-token quantizer (include/b200q.h
b200q_had_quant_rows; reference viditq_quant_layer.py:58-66, quarot_quant_layer.py:55-62, sq_quant_layer.py:55-58).

CPU: the factored plan (colscale, order-K block, 2^w-wide FWHT) is the same operator as `x*mask @ rotation_matrix`.
GPU: the kernel against the float64 restatement of that operator; codes / delta / rowsum bit-exact w.r.t. the kernel's own
rotated activations."""
import pytest
import torch
import torch.nn as nn
from omegaconf import OmegaConf

import fake_backend

SIZES = [256, 384, 512, 768, 1536, 4096, 5120]          # K,w: (1,8) (12,5) (2,8) (12,6) (12,7) (16,8) (20,8)


def _layer(kind, n, cout=64, device="cpu", seed=0):
    from qdiff.quarot.quarot_quant_layer import QuarotQuantizedLinear
    from qdiff.viditq.viditq_quant_layer import ViDiTQuantizedLinear
    from qdiff.smooth_quant.sq_quant_layer import SQQuantizedLinear
    g = torch.Generator().manual_seed(seed)
    fp = nn.Linear(n, cout)
    with torch.no_grad():
        fp.weight.copy_((torch.rand(cout, n, generator=g) * 2 - 1) * 0.1)
        fp.bias.copy_(torch.randn(cout, generator=g) * 0.05)
    fp = fp.to(device)
    extra = {"quarot": {"quarot": {}}, "viditq": {"viditq": {"alpha": 0.6}}, "sq": {"smooth_quant": {"alpha": 0.6}}}[kind]
    cfg = OmegaConf.create({"weight": {"n_bits": 8, "sym": False}, "act": {"n_bits": 8, "sym": True}, **extra})
    cls = {"quarot": QuarotQuantizedLinear, "viditq": ViDiTQuantizedLinear, "sq": SQQuantizedLinear}[kind]
    layer = cls(n, cout, True, None, cfg, fp)
    if kind in ("viditq", "sq"):
        layer.get_channel_mask((torch.rand(n, generator=g) * 4 + 0.1).to(device))
    if kind in ("viditq", "quarot"):
        layer.get_rotation_matrix()
    {"quarot": lambda: layer.update_quantized_weight_rotated(),
     "viditq": lambda: layer.update_quantized_weight_rotated_and_scaled(),
     "sq": lambda: layer.update_quantized_weight_scaled()}[kind]()
    return layer, g


def test_kernel_plan_factorisation():
    from qdiff.quarot.quarot_utils import hadamard_kernel_plan
    want = {256: (1, 8), 384: (12, 5), 512: (2, 8), 768: (12, 6), 1536: (12, 7), 4096: (16, 8), 5120: (20, 8)}
    for n, (K, w) in want.items():
        kp = hadamard_kernel_plan(n)
        assert kp[:2] == (K, w) and (kp[2] is None) == (K == 1)
    assert hadamard_kernel_plan(96) is None            # not a multiple of 128
    assert hadamard_kernel_plan(8960) is None          # order-140 base block: beyond the kernel's K <= 32
    assert hadamard_kernel_plan(13824) is None         # order 108


@pytest.mark.parametrize("kind", ["quarot", "viditq", "sq"])
@pytest.mark.parametrize("n", SIZES)
def test_act_plan_is_the_layer_transform_cpu(monkeypatch, kind, n):
    """plan.colscale / hadK / (K, w) evaluated in float64 == (x * mask) @ rotation_matrix (the reference's forward)."""
    fake_backend.install(monkeypatch)
    layer, g = _layer(kind, n)
    plan = layer._act_plan(torch.device("cpu"))
    assert plan is not None
    x = torch.randn(5, n, generator=g, dtype=torch.float64)
    ref = x
    if kind in ("viditq", "sq"):
        ref = ref * layer.channel_mask.detach().double().reshape(1, -1)
    if kind in ("viditq", "quarot"):
        ref = ref @ layer.rotation_matrix.double()
    got = fake_backend.had_transform(x, plan.colscale, plan.hadK, plan.K, plan.log2w)
    assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6 * float(ref.abs().max()))
    # the layer forward goes through the fused entry and still computes the layer function
    with torch.no_grad():
        y = layer(x.float().unsqueeze(0))[0]
        y_fp = torch.nn.functional.linear(x.float(), layer.fp_module.weight, layer.fp_module.bias)
    cos = float((y.double().flatten() @ y_fp.double().flatten()) / (y.double().norm() * y_fp.double().norm()))
    assert cos > 0.999, cos


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype", [(256, torch.float32), (384, torch.float32), (512, torch.bfloat16), (768, torch.float32),
                                     (1536, torch.float32), (1536, torch.bfloat16), (4096, torch.float32),
                                     (5120, torch.float32), (5120, torch.bfloat16), (3072, torch.float16)])
def test_had_quant_rows_gpu(dev, n, dtype):
    import b200q
    from qdiff.base.quant_layer import ActPlan
    g = torch.Generator().manual_seed(n)
    rows = 137                                               # ragged against every rows-per-CTA choice
    x = (torch.randn(rows, n, generator=g) * (torch.rand(rows, 1, generator=g) * 5 + 0.01)).to(dtype)
    x[3] = 0                                                 # all-zero row: delta floor 1e-6
    x[:, 7] *= 40                                            # outlier channel
    sign = torch.randint(0, 2, (n,), generator=g).float() * 2 - 1
    mask = torch.rand(n, generator=g) * 3 + 0.05
    plan = ActPlan.rotation(n, sign, mask, dev)
    q, d, rs, y = b200q.had_quant_rows(x.to(dev), plan.colscale, plan.hadK, plan.K, plan.log2w, 8, want_y=True)
    ref = fake_backend.had_transform(x.double(), plan.colscale.cpu(), None if plan.hadK is None else plan.hadK.cpu(),
                                     plan.K, plan.log2w)
    y = y.cpu()
    assert float((y.double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max()) * max(1.0, n ** 0.5 / 16)
    # orthogonality: the rotation preserves the row norms of x * mask
    assert torch.allclose(y.double().norm(dim=1), (x.double() * mask.double()).norm(dim=1), rtol=1e-5, atol=1e-6)
    # quantizer part: bit-exact w.r.t. the kernel's own rotated activations
    from oracle import fakequant_oracle as O
    qo, do, _ = O.quant_rows(y, 8, True, True)
    assert torch.equal(q.cpu().float(), qo) and torch.equal(d.cpu(), do.flatten())
    assert torch.equal(rs.cpu(), qo.to(torch.int32).sum(dim=1).to(torch.int32))
    assert float(d[3]) == float(torch.tensor(1e-6, dtype=torch.float32)) and int(q[3].abs().max()) == 0


@pytest.mark.gpu
def test_scale_only_and_errors_gpu(dev):
    import b200q
    from qdiff.base.quant_layer import ActPlan
    from oracle import fakequant_oracle as O
    g = torch.Generator().manual_seed(5)
    x = torch.randn(70, 8960, generator=g).to(torch.bfloat16)
    mask = torch.rand(8960, generator=g) * 2 + 0.1
    plan = ActPlan.scale_only(mask, dev)
    q, d, rs, y = b200q.had_quant_rows(x.to(dev), plan.colscale, None, 1, 0, 8, want_y=True)
    ref = x.float() * mask.reshape(1, -1)
    assert torch.equal(y.cpu(), ref)
    qo, do, _ = O.quant_rows(ref, 8, True, True)
    assert torch.equal(q.cpu().float(), qo) and torch.equal(d.cpu(), do.flatten())
    with pytest.raises(b200q.B200QError):
        b200q.had_quant_rows(x.to(dev)[:, :96].contiguous(), None, None, 1, 0, 8)            # cols % 128 != 0
    with pytest.raises(b200q.B200QError):
        b200q.had_quant_rows(x.to(dev), None, None, 1, 7, 8)                                  # cols != K << w


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["quarot", "viditq", "sq"])
@pytest.mark.parametrize("n", [384, 1536, 5120])
def test_variant_layer_fused_vs_reference_formula_gpu(dev, kind, n):
    """Layer forward through the fused kernel vs the reference's forward formula evaluated with the oracle quantizers on
    the CPU: x*mask, fp64 rotation, fake-quant activations, F.linear on the layer's own fake-quant weight."""
    from oracle import fakequant_oracle as O
    layer, g = _layer(kind, n, cout=96, device=dev)
    assert layer._act_plan(dev) is not None
    x = torch.randn(2, 50, n, generator=g) * 2
    x[..., 11] *= 25
    with torch.no_grad():
        y = layer(x.to(dev)).float().cpu()
    xr = x.reshape(-1, n).double()
    if kind in ("viditq", "sq"):
        xr = xr * layer.channel_mask.detach().double().cpu().reshape(1, -1)
    if kind in ("viditq", "quarot"):
        xr = xr @ layer.rotation_matrix.double().cpu()
    xq = O.fake_quant_rows(xr.float(), 8, True, True)
    ref = torch.nn.functional.linear(xq, layer.weight.data.float().cpu(), layer.bias.detach().float().cpu()).reshape(2, 50, -1)
    rel = float((y - ref).abs().max() / ref.abs().max())
    cos = float((y.double().flatten() @ ref.double().flatten()) / (y.double().norm() * ref.double().norm()))
    assert cos >= 0.99999 and rel <= 5e-3, (cos, rel)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dtype", [(1536, torch.float32), (1536, torch.bfloat16), (1024, torch.float16), (5120, torch.float32),
                                     (5120, torch.bfloat16)])
def test_had_quant_warp_kernel_vs_tile_kernel_gpu(dev, n, dtype):
    """n = K*128 rows take the register-resident warp-per-row kernel, n = 20*256 (the 14B hidden size) its two-warps-per-row
    sibling (csrc/hadamard.cu); the shared-memory tile kernel stays
    selectable: same function, fp32 rounding order differs (base block before / after the Walsh-Hadamard stages)."""
    import b200q
    from qdiff.base.quant_layer import ActPlan
    g = torch.Generator().manual_seed(n + 1)
    rows = 1001
    x = (torch.randn(rows, n, generator=g) * (torch.rand(rows, 1, generator=g) * 4 + 0.1)).to(dtype).to(dev)
    sign = torch.randint(0, 2, (n,), generator=g).float() * 2 - 1
    mask = torch.rand(n, generator=g) * 2 + 0.1
    plan = ActPlan.rotation(n, sign, mask, dev)
    lib = b200q.load()
    try:
        lib.b200q_had_set_mode(0)
        q0, d0, rs0, y0 = b200q.had_quant_rows(x, plan.colscale, plan.hadK, plan.K, plan.log2w, 8, want_y=True)
    finally:
        lib.b200q_had_set_mode(1)
    q1, d1, rs1, y1 = b200q.had_quant_rows(x, plan.colscale, plan.hadK, plan.K, plan.log2w, 8, want_y=True)
    scale = float(y0.abs().max())
    assert float((y0 - y1).abs().max()) <= 4e-6 * scale
    assert torch.allclose(d0, d1, rtol=1e-5)
    assert float((q0.float() - q1.float()).abs().max()) <= 1          # a rounding-order difference can move a code by one step
    assert float((q0 != q1).float().mean()) < 2e-3
    from oracle import fakequant_oracle as O
    qo, do, _ = O.quant_rows(y1.cpu(), 8, True, True)
    assert torch.equal(q1.cpu().float(), qo) and torch.equal(d1.cpu(), do.flatten())
    assert torch.equal(rs1.cpu(), qo.to(torch.int32).sum(dim=1).to(torch.int32))
