"""CPU suite: the oracle restatement (torch + C) against the golden vectors that oracle/gen_golden.py produced by
running the IMPORTED, unmodified reference.  Bit-exact for codes / delta / zero_point / dequantised values."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import build_c
from oracle import fakequant_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name", ["sym8", "asym8", "sym4", "asym4", "ties_sym8"])
def test_dynamic_quantizer(golden_dir, name):
    rec = _load(golden_dir, "dynamic_quantizer.pt")[name]
    q, d, z = O.quant_rows(rec["x"], rec["n_bits"], rec["sym"], dynamic=True)
    assert torch.equal(d, rec["delta"]) and torch.equal(z, rec["zero_point"])
    assert torch.equal(q, rec["codes"])
    assert torch.equal(O.dequant_rows(q, d, z), rec["dequant"])


@pytest.mark.parametrize("name", ["asym8", "asym4"])
def test_static_quantizer(golden_dir, name):
    rec = _load(golden_dir, "static_quantizer.pt")[name]
    q, d, z = O.quant_rows(rec["w"], rec["n_bits"], False, dynamic=False)
    assert torch.equal(d, rec["delta"]) and torch.equal(z, rec["zero_point"]) and torch.equal(q, rec["codes"])
    assert torch.equal(O.dequant_rows(q, d, z), rec["dequant"])


@pytest.mark.parametrize("name", ["w8a8", "w4a8"])
def test_quantized_linear(golden_dir, name):
    rec = _load(golden_dir, "quantized_linear.pt")[name]
    y = O.quantized_linear_fake(rec["x"], rec["weight"], rec["bias"], w_bits=rec["w_bits"])
    assert torch.equal(y, rec["y"])
    # the real-integer algebra the CUDA path implements agrees with the fake-quant result
    yi, acc, qa, da, zpa, qw, dw, zpw = O.quantized_linear_int(rec["x"], rec["weight"], rec["bias"], w_bits=rec["w_bits"])
    assert torch.equal(dw, rec["w_delta"]) and torch.equal(zpw, rec["w_zero_point"])
    assert torch.equal(O.dequant_rows(qw, dw, zpw), rec["w_dequant"])
    assert float((yi - rec["y"]).abs().max()) <= 2e-5 * float(rec["y"].abs().max() + 1)
    cos = float((yi.flatten().double() @ rec["y"].flatten().double()) / (yi.norm().double() * rec["y"].norm().double()))
    assert cos >= 0.999999


def test_mixed_precision(golden_dir):
    rec = _load(golden_dir, "mixed_precision.pt")
    for i, bits in enumerate((4, 8)):
        q, d, z = O.quant_rows(rec["w"], bits, False, dynamic=False)
        g = rec[f"static_i{i}"]
        assert torch.equal(d, g["delta"]) and torch.equal(z, g["zero_point"])
        assert torch.equal(O.dequant_rows(q, d, z), g["dequant"])
        assert torch.equal(d, g["delta_list"][i]) and torch.equal(z, g["zero_point_list"][i])
        # after bitwidth_refactor(1-i) the reference uses the other entry's parameters
        other = (8, 4)[i]
        qo, do, zo = O.quant_rows(rec["w"], other, False, dynamic=False)
        assert torch.equal(O.dequant_rows(qo, do, zo), rec[f"static_i{i}_refactored"]["dequant"])
        gd = rec[f"dynamic_i{i}"]
        assert torch.equal(O.fake_quant_rows(gd["x"], bits, True, True), gd["dequant"])


def test_calibration(golden_dir):
    rec = _load(golden_dir, "calibration.pt")
    per_call = [O.calib_absmax(c) for c in rec["calls"]]
    for a, b in zip(per_call, rec["per_call"]):
        assert torch.equal(a, b)
    assert torch.equal(O.calib_merge(per_call), rec["merged"])


def test_quantized_attention(golden_dir):
    rec = _load(golden_dir, "quant_attention.pt")
    out, info = O.quantized_attention_fake(rec["q"], rec["k"], rec["v"], p_sym=False)
    assert torch.equal(info["dq"].flatten(), rec["q_delta"].flatten())
    assert torch.equal(info["dk"].flatten(), rec["k_delta"].flatten())
    assert torch.equal(info["dv"].flatten(), rec["v_delta"].flatten())
    assert torch.equal(info["dp"].flatten(), rec["p_delta"].flatten())
    assert torch.equal(info["zp"].flatten(), rec["p_zero_point"].flatten())
    assert torch.equal(out, rec["out"])


def test_quantized_attention_rowstep(golden_dir):
    """Fused-kernel attention semantics (P quantized on the unsigned grid of forward_with_quant_params with one step per
    query row): the oracle reproduces the imported reference quantizers bit-for-bit (oracle/gen_golden_attn.py)."""
    rec = _load(golden_dir, "quant_attention_rowstep.pt")
    out, info = O.quantized_attention_rowstep(rec["q"], rec["k"], rec["v"])
    assert torch.equal(info["dq"].flatten(), rec["q_delta"].flatten())
    assert torch.equal(info["dk"].flatten(), rec["k_delta"].flatten())
    assert torch.equal(info["dv"].flatten(), rec["v_delta"].flatten())
    assert torch.equal(info["attn"], rec["attn"])
    assert torch.equal(info["p_dequant"], rec["attn_quant"])
    assert torch.equal(out, rec["out"])
    assert info["p_codes"].max() == 255 and info["p_codes"].min() == 0


# ---- C restatement ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def clib():
    lib = ctypes.CDLL(build_c.build())
    lib.div_hoisted_check.restype = ctypes.c_long
    lib.div_hoisted_check.argtypes = [ctypes.c_long, ctypes.c_uint64]
    return lib


@pytest.mark.parametrize("name", ["sym8", "asym8", "sym4", "asym4", "ties_sym8"])
def test_c_oracle_quantizer(golden_dir, clib, name):
    rec = _load(golden_dir, "dynamic_quantizer.pt")[name]
    x = np.ascontiguousarray(rec["x"].numpy())
    rows, cols = x.shape
    codes = np.empty_like(x); delta = np.empty(rows, np.float32); zp = np.empty(rows, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    clib.quant_rows_f32(p(x), ctypes.c_long(rows), ctypes.c_long(cols), rec["n_bits"], int(rec["sym"]), 1, p(codes),
                        p(delta), p(zp))
    assert np.array_equal(delta, rec["delta"].flatten().numpy())
    assert np.array_equal(zp, rec["zero_point"].flatten().numpy())
    assert np.array_equal(codes, rec["codes"].numpy())


def test_c_oracle_int_gemm(clib):
    g = torch.Generator().manual_seed(0)
    a = torch.randint(-128, 128, (37, 200), generator=g, dtype=torch.int32).to(torch.int8)
    w = torch.randint(-128, 128, (29, 200), generator=g, dtype=torch.int32).to(torch.int8)
    acc = np.empty((37, 29), np.int32)
    p = lambda t: t.ctypes.data_as(ctypes.c_void_p)
    an, wn = a.numpy(), w.numpy()
    clib.gemm_i8_i32(p(an), p(wn), ctypes.c_long(37), ctypes.c_long(29), ctypes.c_long(200), p(acc))
    assert np.array_equal(acc, O.int_accumulators(a, w).numpy())


def test_hoisted_division_is_ieee_division(clib):
    """The CUDA quantizer replaces x/delta by a reciprocal + two exact-remainder corrections (FFMA2); this is the
    same sequence on the CPU (fmaf) against true IEEE division on 16M random / near-tie cases: zero mismatches."""
    assert clib.div_hoisted_check(2_000_000, 12345) == 0


def test_c_oracle_attention_rowstep(golden_dir, clib):
    """The integer form of the fused attention (C, double-precision softmax) against the golden vectors of the imported
    reference quantizers: codes within one step of the reference's fp32-softmax codes (ties only), outputs to 1e-3."""
    rec = _load(golden_dir, "quant_attention_rowstep.pt")
    _, info = O.quantized_attention_rowstep(rec["q"], rec["k"], rec["v"])
    B, H, Lq, hd = rec["q"].shape
    Lk = rec["k"].shape[2]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    step = rec["attn"].max(dim=-1, keepdim=True)[0] / 255.0
    ref_codes = torch.round(rec["attn_quant"] / step)[0]
    for h in range(H):
        qq = np.ascontiguousarray(info["qq"][0, h].numpy().astype(np.int8))
        kq = np.ascontiguousarray(info["kq"][0, h].numpy().astype(np.int8))
        vt = np.ascontiguousarray(info["vq"][0, h].numpy().astype(np.int8))            # [hd, Lk]
        dq = np.ascontiguousarray(info["dq"][0, h].numpy()); dk = np.ascontiguousarray(info["dk"][0, h].numpy())
        dv = np.ascontiguousarray(info["dv"][0, h].numpy())
        codes = np.empty((Lq, Lk), np.uint8); acc = np.empty((Lq, hd), np.int64); out = np.empty((Lq, hd), np.float32)
        clib.attn_rowstep_i8(p(qq), p(kq), p(vt), p(dq), p(dk), p(dv), ctypes.c_long(Lq), ctypes.c_long(Lk), ctypes.c_long(hd),
                             ctypes.c_double(hd ** -0.5), p(codes), p(acc), p(out), None)
        d = np.abs(codes.astype(np.int32) - ref_codes[h].numpy().astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.01
        assert np.array_equal(acc, codes.astype(np.int64) @ vt.astype(np.int64).T)
        ref_out = rec["out"][0, h].numpy()
        assert np.abs(out - ref_out).max() <= 1e-3 * np.abs(ref_out).max() + 1e-6
