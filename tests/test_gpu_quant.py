"""-m gpu parity tests: CUDA quantizer / calibration kernels (through the C ABI) vs the CPU oracle and the
golden vectors produced by the imported reference.  Codes, delta, zero_point: BIT-EXACT."""
import os

import pytest
import torch

import b200q
from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu


def _act(rows, cols, seed, outliers=True, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, cols, generator=g) * (torch.rand(rows, 1, generator=g) * 4 + 0.05)
    if outliers and cols >= 100:
        idx = torch.randint(0, cols, (max(1, cols // 1000),), generator=g)
        x[:, idx] *= 50.0
    return x.to(dtype)


def _check_quant(x_cpu, dev, n_bits, sym, dynamic):
    q, d, z, rs = b200q.quant_rows(x_cpu.to(dev), n_bits, sym, dynamic)
    qo, do, zo = O.quant_rows(x_cpu.float(), n_bits, sym, dynamic)
    assert torch.equal(d.cpu(), do.flatten()), "delta not bit-exact"
    assert torch.equal(z.cpu(), zo.flatten()), "zero_point not bit-exact"
    qo_i8 = qo.clamp(-128, 127).to(torch.int8)
    assert torch.equal(q.cpu(), qo_i8), f"codes differ in {(q.cpu() != qo_i8).sum().item()} places"
    assert torch.equal(rs.cpu(), qo_i8.to(torch.int32).sum(dim=1).to(torch.int32))


@pytest.mark.parametrize("name", ["sym8", "asym8", "sym4", "asym4", "ties_sym8"])
def test_dynamic_quantizer_golden(dev, golden_dir, name):
    rec = torch.load(os.path.join(golden_dir, "dynamic_quantizer.pt"))[name]
    q, d, z, _ = b200q.quant_rows(rec["x"].to(dev), rec["n_bits"], rec["sym"], True)
    assert torch.equal(d.cpu(), rec["delta"].flatten())
    assert torch.equal(z.cpu(), rec["zero_point"].flatten())
    assert torch.equal(q.cpu().float(), rec["codes"].clamp(-128, 127))
    deq = b200q.dequant_rows(q, d, z)
    assert torch.equal(deq.cpu(), rec["dequant"])


@pytest.mark.parametrize("name", ["asym8", "asym4"])
def test_static_quantizer_golden(dev, golden_dir, name):
    rec = torch.load(os.path.join(golden_dir, "static_quantizer.pt"))[name]
    q, d, z, _ = b200q.quant_rows(rec["w"].to(dev), rec["n_bits"], False, False)
    assert torch.equal(d.cpu(), rec["delta"].flatten())
    assert torch.equal(z.cpu(), rec["zero_point"].flatten())
    assert torch.equal(q.cpu().float(), rec["codes"])
    q2, _ = b200q.quant_rows_static(rec["w"].to(dev), d, z, rec["n_bits"], False)
    assert torch.equal(q2.cpu().float(), rec["codes"])
    assert torch.equal(b200q.dequant_rows(q, d, z).cpu(), rec["dequant"])


@pytest.mark.parametrize("rows,cols", [(257, 1536), (64, 5120), (33, 8960), (17, 13824), (5, 128), (3, 40000),
                                       (9, 1000), (7, 37), (1, 8), (300, 256)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_dynamic_sym8_shapes(dev, rows, cols, dtype):
    _check_quant(_act(rows, cols, rows * 7 + cols, dtype=dtype), dev, 8, True, True)


@pytest.mark.parametrize("n_bits,sym,dynamic", [(8, False, True), (4, True, True), (4, False, True),
                                                (8, False, False), (4, False, False), (8, True, False), (6, True, True)])
@pytest.mark.parametrize("cols", [1536, 8960, 100])
def test_quantizer_modes(dev, n_bits, sym, dynamic, cols):
    x = _act(48, cols, 11 + cols)
    x[3] = x[3].abs()
    x[4] = -x[4].abs()
    _check_quant(x, dev, n_bits, sym, dynamic)


def test_eps_floor_and_zero_rows(dev):
    x = _act(16, 1536, 5)
    x[2] = 0.0
    x[3] = x[3] * 1e-9
    _check_quant(x, dev, 8, True, True)


def test_ties_round_half_even(dev):
    # amax = 63.5 -> delta = 0.5 exactly; x on exact .5 multiples -> every quotient is k or k+.5
    x = torch.zeros(8, 256)
    for r in range(8):
        x[r] = (torch.arange(256).float() - 127) * 0.25
        x[r, 0] = -63.5
        x[r, 255] = 63.5
    _check_quant(x, dev, 8, True, True)


def test_strided_input_and_empty(dev):
    big = _act(40, 4096, 3).to(dev)
    view = big[:, 512:2048]                       # ldx = 4096, cols = 1536
    q, d, z, rs = b200q.quant_rows(view, 8, True, True)
    qo, do, _ = O.quant_rows(view.cpu(), 8, True, True)
    assert torch.equal(q.cpu().float(), qo) and torch.equal(d.cpu(), do.flatten())
    e = torch.empty(0, 1536, device=dev)
    q, d, z, rs = b200q.quant_rows(e, 8, True, True)
    assert q.shape == (0, 1536)


def test_quant_full_size_properties(dev):
    """BASELINE config sizes: size-independent properties instead of a CPU re-computation."""
    L, D = 32760, 1536
    x = torch.randn(L, D, device=dev) * 3
    q, d, z, rs = b200q.quant_rows(x, 8, True, True)
    amax = x.abs().amax(dim=1)
    assert torch.equal(d, torch.div(amax, torch.full_like(amax, 127.0)))   # tensor/tensor = IEEE division (tensor/scalar multiplies by 1/127)
    assert int(q.abs().amax()) == 127                          # every row hits +-127 at its abs-max
    assert torch.equal(q.abs().amax(dim=1).int(), torch.full((L,), 127, device=dev, dtype=torch.int32))
    assert torch.equal(rs, q.sum(dim=1, dtype=torch.int32))
    deq = b200q.dequant_rows(q, d, z)
    assert float(((deq - x).abs() / d[:, None]).max()) <= 0.5 + 1e-3   # error bounded by half a step
    # idempotence: quantizing the dequantised tensor reproduces the codes
    q2, d2, _, _ = b200q.quant_rows(deq, 8, True, True)
    assert torch.equal(q2, q)
    # a random 64-row sample against the CPU oracle
    idx = torch.randint(0, L, (64,))
    qo, do, _ = O.quant_rows(x[idx.to(dev)].cpu(), 8, True, True)
    assert torch.equal(q[idx.to(dev)].cpu().float(), qo)


# ---- calibration -----------------------------------------------------------------------------
def test_calibration_golden(dev, golden_dir):
    rec = torch.load(os.path.join(golden_dir, "calibration.pt"))
    C = rec["calls"][0].shape[-1]
    absmax = torch.zeros(C, device=dev)
    for i, call in enumerate(rec["calls"]):
        one = torch.zeros(C, device=dev)
        b200q.calib_update(call.reshape(-1, C).to(dev), one)
        assert torch.equal(one.cpu(), rec["per_call"][i])
        b200q.calib_update(call.reshape(-1, C).to(dev), absmax)
    assert torch.equal(absmax.cpu(), rec["merged"])


@pytest.mark.parametrize("rows,cols", [(4097, 1536), (1000, 8960), (31, 5120), (513, 100), (7, 13), (1, 1536)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_calibration_shapes(dev, rows, cols, dtype):
    x = _act(rows, cols, rows + cols, dtype=dtype)
    absmax = torch.zeros(cols, device=dev)
    mn = torch.full((cols,), float("inf"), device=dev)
    mx = torch.full((cols,), float("-inf"), device=dev)
    b200q.calib_update(x.to(dev), absmax, mn, mx)
    assert torch.equal(absmax.cpu(), O.calib_absmax(x.float()))
    lo, hi = O.calib_minmax(x.float())
    assert torch.equal(mn.cpu(), lo) and torch.equal(mx.cpu(), hi)
    # sharding the rows (sequence parallel) and accumulating gives the identical statistic
    a2 = torch.zeros(cols, device=dev)
    for chunk in torch.chunk(x.to(dev), 4, dim=0):
        if chunk.shape[0]:
            b200q.calib_update(chunk, a2)
    assert torch.equal(a2, absmax)


def test_calibration_full_size(dev):
    L, D = 32760, 1536
    x = torch.randn(L, D, device=dev)
    a = torch.zeros(D, device=dev)
    b200q.calib_update(x, a)
    assert torch.equal(a, x.abs().amax(dim=0))


# ---- fused LN + modulate + quant, gate residual ---------------------------------------------------
@pytest.mark.parametrize("rows,cols", [(200, 1536), (40, 5120), (64, 128)])
@pytest.mark.parametrize("affine,mod", [(False, True), (True, False), (False, False)])
def test_ln_mod_quant(dev, rows, cols, affine, mod):
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g) * 2 + 0.3
    w = torch.randn(cols, generator=g) * 0.2 + 1 if affine else None
    b = torch.randn(cols, generator=g) * 0.1 if affine else None
    shift = torch.randn(cols, generator=g) * 0.1 if mod else None
    scale = torch.randn(cols, generator=g) * 0.1 if mod else None
    ref = O.layer_norm(x, w, b, 1e-6)
    if mod:
        ref = ref * (1 + scale) + shift
    mv = lambda t: None if t is None else t.to(dev)
    q, d, rs, y = b200q.ln_mod_quant(x.to(dev), 1e-6, mv(w), mv(b), mv(shift), mv(scale), 8, True, True,
                                     y_dtype=torch.float32)
    assert torch.allclose(y.cpu(), ref, rtol=2e-5, atol=2e-5)
    # the quantizer stage is exact given the kernel's own y
    qo, do, _ = O.quant_rows(y.cpu(), 8, True, True)
    assert torch.equal(d.cpu(), do.flatten())
    assert torch.equal(q.cpu().float(), qo)
    assert torch.equal(rs.cpu(), qo.sum(dim=1).to(torch.int32))
    # and within one code of the oracle's LN
    qr, _, _ = O.quant_rows(ref, 8, True, True)
    diff = (q.cpu().float() - qr).abs()
    assert diff.max() <= 1 and (diff > 0).float().mean() < 2e-3


@pytest.mark.parametrize("ydt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("rows,cols", [(100, 1536), (7, 30)])
def test_gate_residual(dev, ydt, rows, cols):
    g = torch.Generator().manual_seed(1)
    y = torch.randn(rows, cols, generator=g).to(ydt)
    res = torch.randn(rows, cols, generator=g)
    gate = torch.randn(cols, generator=g)
    out = b200q.gate_residual(y.to(dev), res.clone().to(dev), gate.to(dev))
    assert torch.equal(out.cpu(), res + y.float() * gate)
    out = b200q.gate_residual(y.to(dev), res.clone().to(dev), None)
    assert torch.equal(out.cpu(), res + y.float())


def test_mixed_precision_quantizers_golden_gpu(dev, golden_dir):
    """MixedPrecisionStaticQuantizer / MixedPrecisionDynamicQuantizer of the mirror on libb200q vs the outputs of the
    imported reference classes (mixed_precision_quantizer.py:56-186; golden/mixed_precision.pt): delta / zero_point lists,
    the selected entry, the dequantised tensor before and after bitwidth_refactor - all bit-exact."""
    import os
    from omegaconf import OmegaConf
    from qdiff.base.mixed_precision_quantizer import MixedPrecisionDynamicQuantizer, MixedPrecisionStaticQuantizer
    rec = torch.load(os.path.join(golden_dir, "mixed_precision.pt"))
    w = rec["w"].to(dev)
    for i in (0, 1):
        qz = MixedPrecisionStaticQuantizer(OmegaConf.create({"n_bits": [4, 8], "sym": False, "i_bitwidth": i}))
        deq = qz.forward(w.clone())
        g = rec[f"static_i{i}"]
        assert torch.equal(qz.delta.cpu(), g["delta"]) and torch.equal(qz.zero_point.cpu(), g["zero_point"])
        assert torch.equal(qz.delta_list.cpu(), g["delta_list"]) and torch.equal(qz.zero_point_list.cpu(), g["zero_point_list"])
        assert torch.equal(deq.cpu(), g["dequant"])
        qz.init_done = True
        qz.bitwidth_refactor(1 - i)
        gr = rec[f"static_i{i}_refactored"]
        assert torch.equal(qz.delta.cpu(), gr["delta"]) and torch.equal(qz.forward(w.clone()).cpu(), gr["dequant"])
        assert qz.n_bits == (8, 4)[i]
        gd = rec[f"dynamic_i{i}"]
        dz = MixedPrecisionDynamicQuantizer(OmegaConf.create({"n_bits": [4, 8], "sym": True, "i_bitwidth": i}))
        assert torch.equal(dz.forward(gd["x"].to(dev)).cpu(), gd["dequant"]) and torch.equal(dz.delta.cpu(), gd["delta"])
        # the real-integer entry the quantized linear uses gives the same codes
        q, d, _, rs = dz.quantize_int8(gd["x"].to(dev))
        assert torch.equal((q.float() * d.unsqueeze(1)).cpu(), gd["dequant"]) and torch.equal(rs.cpu(), q.cpu().to(torch.int32).sum(1).to(torch.int32))


def test_static_quantizer_running_statistics_and_fp32_params_gpu(dev):
    """base_quantizer.py:76-88: x_max / x_min are running statistics over init_quant_params calls while delta follows the
    current rows; 16-bit weights: the integer path keeps the exact fp32 parameters behind the rounded buffers."""
    from omegaconf import OmegaConf
    from qdiff.base.base_quantizer import StaticQuantizer
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(6, 64, generator=g).to(dev), (torch.randn(6, 64, generator=g) * 3).to(dev)
    qz = StaticQuantizer(OmegaConf.create({"n_bits": 8, "sym": False}))
    qz.init_quant_params(a)
    qz.init_quant_params(b)
    assert torch.equal(qz.x_max, torch.max(a.max(1)[0].clamp_min(0), b.max(1)[0].clamp_min(0)))
    assert torch.equal(qz.x_min, torch.min(a.min(1)[0].clamp_max(0), b.min(1)[0].clamp_max(0)))
    _, d, z = O.quant_rows(b.cpu(), 8, False, dynamic=False)
    assert torch.equal(qz.delta.cpu(), d)
    wb = (torch.randn(8, 128, generator=g) * 0.1).to(torch.bfloat16).to(dev)
    qb = StaticQuantizer(OmegaConf.create({"n_bits": 8, "sym": False}))
    deq = qb.forward(wb)
    qo, do, zo = O.quant_rows(wb.float().cpu(), 8, False, dynamic=False)
    d32, z32 = qb.params_f32(dev)
    assert torch.equal(d32.cpu(), do.flatten()) and torch.equal(z32.cpu(), zo.flatten())
    qb.init_done = True
    assert torch.equal(qb.quantize_int8(wb).cpu().float(), qo.clamp(-128, 127))      # codes from the exact parameters
    assert deq.dtype == torch.bfloat16 and qb.delta.dtype == torch.bfloat16
