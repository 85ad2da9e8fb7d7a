"""-m gpu parity tests: tcgen05 int8 GEMM through the C ABI.  int32 accumulators BIT-EXACT vs the oracle;
dequantised outputs within the stated tolerance (cosine >= 0.999, max rel err stated per test)."""
import os

import pytest
import torch

import b200q
from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu


def _codes(rows, cols, seed, lo=-127, hi=127):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, (rows, cols), generator=g, dtype=torch.int32).to(torch.int8)


SHAPES = [(128, 256, 128), (256, 512, 256), (384, 256, 1536), (130, 272, 144), (1, 8, 16), (515, 1536, 1536),
          (100, 100, 4096), (128, 8960, 320), (777, 300, 208),
          # the Wan-14B GEMM dims of SURVEY §8d (reduced M): K = 13824 (ffn.2), N = 13824 (ffn.0), N = K = 5120
          (200, 5120, 13824), (150, 13824, 5120), (260, 5120, 5120), (140, 15360, 5120), (200, 1536, 8960)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_int32_accumulators_bit_exact(dev, M, N, K):
    qa, qw = _codes(M, K, M + K), _codes(N, K, N + K + 1, -128, 127)
    acc = b200q.gemm_w8a8(qa.to(dev), qw.to(dev), out_dtype=torch.int32)
    assert torch.equal(acc.cpu(), O.int_accumulators(qa, qw))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("out_dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 8e-3), (torch.float16, 1e-3)])
@pytest.mark.parametrize("M,N,K", [(300, 520, 1536), (64, 256, 8960), (129, 40, 144)])
def test_dequant_epilogue(dev, M, N, K, out_dtype, tol):
    """Distribution of ViDiT-Q/kernels/bench/bench_gemm.py:4-16 (codes in [-80,80), positive scales, zp in [-10,10))."""
    g = torch.Generator().manual_seed(M * 3 + N)
    qa, qw = _codes(M, K, 1, -80, 79), _codes(N, K, 2, -80, 79)
    da = torch.rand(M, generator=g) * 0.01 + 0.005
    dw = torch.rand(N, generator=g) * 0.1 + 0.1
    zp = torch.randint(-10, 10, (N,), generator=g).float()
    bias = torch.rand(N, generator=g) * 200
    rs = qa.to(torch.int32).sum(dim=1).to(torch.int32)
    acc = O.int_accumulators(qa, qw).double()
    ref = (da.double()[:, None] * dw.double()[None, :] * (acc + zp.double()[None, :] * rs.double()[:, None])
           + bias.double()[None, :])
    out = b200q.gemm_w8a8(qa.to(dev), qw.to(dev), da.to(dev), dw.to(dev), zp.to(dev), rs.to(dev), bias.to(dev),
                          out_dtype=out_dtype)
    got = out.cpu().double()
    rel = float(((got - ref).abs() / (ref.abs() + 1.0)).max())
    assert rel <= tol, rel
    assert _cos(got, ref) >= 0.99999
    # symmetric weights, no bias
    out = b200q.gemm_w8a8(qa.to(dev), qw.to(dev), da.to(dev), dw.to(dev), out_dtype=out_dtype)
    ref2 = da.double()[:, None] * dw.double()[None, :] * acc
    rel = float(((out.cpu().double() - ref2).abs() / (ref2.abs() + 1.0)).max())
    assert rel <= tol, rel


@pytest.mark.parametrize("name", ["w8a8"])
def test_quantized_linear_golden(dev, golden_dir, name):
    """End to end vs the imported reference's QuantizedLinear.forward (quant_layer.py:57-74): quantizer kernels +
    GEMM kernel reproduce its fp32 fake-quant output.  Tolerance: max |err| <= 2e-5 * (|y| + 1) in fp32."""
    rec = torch.load(os.path.join(golden_dir, "quantized_linear.pt"))[name]
    x = rec["x"].reshape(-1, rec["x"].shape[-1]).to(dev)
    qw, dw, zw, _ = b200q.quant_rows(rec["weight"].to(dev), rec["w_bits"], False, False, want_rowsum=False)
    assert torch.equal(dw.cpu(), rec["w_delta"].flatten()) and torch.equal(zw.cpu(), rec["w_zero_point"].flatten())
    qa, da, _, rs = b200q.quant_rows(x, 8, True, True)
    assert torch.equal(da.cpu(), rec["a_delta"].flatten())
    K = x.shape[1]
    Kp = (K + 15) // 16 * 16                                   # TMA needs 16-byte row pitch: pad K with zero codes
    qa_p = torch.zeros(qa.shape[0], Kp, dtype=torch.int8, device=dev); qa_p[:, :K] = qa
    qw_p = torch.zeros(qw.shape[0], Kp, dtype=torch.int8, device=dev); qw_p[:, :K] = qw
    y = b200q.gemm_w8a8(qa_p[:, :K], qw_p[:, :K], da, dw, zw, rs, rec["bias"].to(dev), out_dtype=torch.float32)
    ref = rec["y"].reshape(-1, rec["y"].shape[-1])
    err = (y.cpu() - ref).abs() / (ref.abs() + 1.0)
    assert float(err.max()) <= 2e-5, float(err.max())
    assert _cos(y.cpu(), ref) >= 0.999999


def test_gelu_epilogue(dev):
    M, N, K = 200, 512, 256
    g = torch.Generator().manual_seed(9)
    qa, qw = _codes(M, K, 3), _codes(N, K, 4)
    da = torch.rand(M, generator=g) * 2e-4 + 1e-5
    dw = torch.rand(N, generator=g) * 1e-2 + 1e-3
    bias = torch.randn(N, generator=g) * 0.1
    pre = da[:, None].double() * dw[None, :].double() * O.int_accumulators(qa, qw).double() + bias.double()
    ref = torch.nn.functional.gelu(pre.float(), approximate="tanh")
    out = b200q.gemm_w8a8(qa.to(dev), qw.to(dev), da.to(dev), dw.to(dev), None, None, bias.to(dev),
                          out_dtype=torch.float32, epilogue=b200q.EPI_GELU_TANH)
    assert float((out.cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max() + 1)   # tanh.approx.f32: ~2^-11 rel
    assert _cos(out.cpu(), ref) >= 0.99999


def test_gate_residual_epilogue(dev):
    M, N, K = 333, 768, 512
    g = torch.Generator().manual_seed(10)
    qa, qw = _codes(M, K, 5), _codes(N, K, 6)
    da = torch.rand(M, generator=g) * 2e-4 + 1e-5
    dw = torch.rand(N, generator=g) * 1e-2 + 1e-3
    zp = torch.randint(-3, 4, (N,), generator=g).float()
    bias = torch.randn(N, generator=g) * 0.1
    gate = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    rs = qa.to(torch.int32).sum(dim=1).to(torch.int32)
    acc = O.int_accumulators(qa, qw).double() + zp.double()[None, :] * rs.double()[:, None]
    y = da[:, None].double() * dw[None, :].double() * acc + bias.double()
    ref = res.double() + y * gate.double()
    x = res.clone().to(dev)
    out = b200q.gemm_w8a8(qa.to(dev), qw.to(dev), da.to(dev), dw.to(dev), zp.to(dev), rs.to(dev), bias.to(dev),
                          epilogue=b200q.EPI_GATE_RESIDUAL, residual=x, gate=gate.to(dev))
    assert out.data_ptr() == x.data_ptr()
    assert float((out.cpu().double() - ref).abs().max()) <= 1e-4


def test_full_size_linearity_and_checksum(dev):
    """BASELINE config 1 shape (32760 x 1536 x 1536): properties instead of a CPU recomputation —
    column checksum of the accumulators equals (sum_m qa) @ qw^T computed in int64 on a reduced problem,
    and acc(qa1 + qa2) == acc(qa1) + acc(qa2)."""
    L, D = 32760, 1536
    qa1 = torch.randint(-60, 61, (L, D), device=dev, dtype=torch.int8)
    qa2 = torch.randint(-60, 61, (L, D), device=dev, dtype=torch.int8)
    qw = torch.randint(-128, 128, (D, D), device=dev, dtype=torch.int8)
    a1 = b200q.gemm_w8a8(qa1, qw, out_dtype=torch.int32)
    a2 = b200q.gemm_w8a8(qa2, qw, out_dtype=torch.int32)
    a12 = b200q.gemm_w8a8(qa1 + qa2, qw, out_dtype=torch.int32)
    assert torch.equal(a12, a1 + a2)
    colsum = a1.sum(dim=0, dtype=torch.int64)
    ref = (qa1.sum(dim=0, dtype=torch.int64).double()[None, :] @ qw.double().t()).flatten()
    assert torch.equal(colsum.double(), ref)
    idx = torch.randint(0, L, (48,), device=dev)
    assert torch.equal(a1[idx].cpu(), O.int_accumulators(qa1[idx].cpu(), qw.cpu()))


def test_bad_arguments_raise(dev):
    qa = torch.zeros(8, 24, dtype=torch.int8, device=dev)      # lda = 24 is not a multiple of 16
    qw = torch.zeros(8, 24, dtype=torch.int8, device=dev)
    with pytest.raises(b200q.B200QError):
        b200q.gemm_w8a8(qa, qw, out_dtype=torch.int32)
    with pytest.raises(b200q.B200QError):
        b200q.quant_rows(torch.zeros(4, 4), 8, True, True)     # CPU tensor: no fallback


# ---- W4A8 ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(256, 512, 256), (300, 520, 1536), (64, 256, 8960), (130, 40, 144), (77, 264, 208)])
@pytest.mark.parametrize("with_zp", [True, False])
def test_w4a8_gemm(dev, M, N, K, with_zp):
    g = torch.Generator().manual_seed(M + N + K)
    qa = _codes(M, K, 21)
    qw = _codes(N, K, 22, -8, 7)
    da = torch.rand(M, generator=g) * 0.01 + 0.005
    dw = torch.rand(N, generator=g) * 0.1 + 0.1
    zp = torch.randint(-3, 4, (N,), generator=g).float() if with_zp else None
    bias = torch.randn(N, generator=g)
    rs = qa.to(torch.int32).sum(dim=1).to(torch.int32)
    acc = O.int_accumulators(qa, qw).double()
    if with_zp:
        acc = acc + zp.double()[None, :] * rs.double()[:, None]
    ref = da.double()[:, None] * dw.double()[None, :] * acc + bias.double()[None, :]
    packed = b200q.pack_w4(qw.to(dev))
    out = b200q.gemm_w4a8(qa.to(dev), packed, K, da.to(dev), dw.to(dev), None if zp is None else zp.to(dev), rs.to(dev),
                          bias.to(dev), out_dtype=torch.float32)
    rel = float(((out.cpu().double() - ref).abs() / (ref.abs() + 1.0)).max())
    assert rel <= 2e-5, rel          # exact integer accumulation: only the fp32 epilogue rounds


def test_pack_w4_saturates_out_of_range_codes(dev):
    """The asymmetric 4-bit quantizer can emit +8 on an exact double rounding tie; the packer stores it as +7 (and
    anything below -8 as -8) instead of letting the nibble wrap to the opposite sign."""
    M, N, K = 64, 32, 128
    qa = _codes(M, K, 5)
    qw = _codes(N, K, 6, -8, 7)
    qw[3, 10], qw[7, 0], qw[9, 127] = 8, 9, -9
    clipped = qw.clamp(-8, 7)
    da, dw = torch.ones(M), torch.ones(N)
    rs = qa.to(torch.int32).sum(dim=1).to(torch.int32)
    out = b200q.gemm_w4a8(qa.to(dev), b200q.pack_w4(qw.to(dev)), K, da.to(dev), dw.to(dev), None, rs.to(dev), None,
                          out_dtype=torch.float32)
    assert torch.equal(out.cpu().double(), O.int_accumulators(qa, clipped).double())


def test_w4a8_quantized_linear_golden(dev, golden_dir):
    """W4 asym weights / A8 sym activations end to end vs the imported reference QuantizedLinear (fp32 fake-quant)."""
    rec = torch.load(os.path.join(golden_dir, "quantized_linear.pt"))["w4a8"]
    x = rec["x"].reshape(-1, rec["x"].shape[-1]).to(dev)
    qw, dw, zw, _ = b200q.quant_rows(rec["weight"].to(dev), 4, False, False, want_rowsum=False)
    assert torch.equal(dw.cpu(), rec["w_delta"].flatten()) and torch.equal(zw.cpu(), rec["w_zero_point"].flatten())
    assert int(qw.min()) >= -8 and int(qw.max()) <= 7
    qa, da, _, rs = b200q.quant_rows(x, 8, True, True)
    K = x.shape[1]
    Kp = (K + 15) // 16 * 16
    qa_p = torch.zeros(qa.shape[0], Kp, dtype=torch.int8, device=dev); qa_p[:, :K] = qa
    y = b200q.gemm_w4a8(qa_p[:, :K], b200q.pack_w4(qw), K, da, dw, zw, rs, rec["bias"].to(dev), out_dtype=torch.float32)
    ref = rec["y"].reshape(-1, rec["y"].shape[-1])
    err = (y.cpu() - ref).abs() / (ref.abs() + 1.0)
    assert float(err.max()) <= 2e-5, float(err.max())


def test_w4a8_gelu_and_gate_epilogues(dev):
    M, N, K = 200, 512, 384
    g = torch.Generator().manual_seed(5)
    qa, qw = _codes(M, K, 31), _codes(N, K, 32, -8, 7)
    da = torch.rand(M, generator=g) * 2e-3 + 1e-4
    dw = torch.rand(N, generator=g) * 1e-2 + 1e-3
    zp = torch.randint(-2, 3, (N,), generator=g).float()
    rs = qa.to(torch.int32).sum(dim=1).to(torch.int32)
    y = da[:, None].double() * dw[None, :].double() * (O.int_accumulators(qa, qw).double() + zp.double()[None, :] * rs.double()[:, None])
    packed = b200q.pack_w4(qw.to(dev))
    out = b200q.gemm_w4a8(qa.to(dev), packed, K, da.to(dev), dw.to(dev), zp.to(dev), rs.to(dev), None,
                          out_dtype=torch.bfloat16, epilogue=b200q.EPI_GELU_TANH)
    ref = torch.nn.functional.gelu(y.float(), approximate="tanh")
    assert _cos(out.cpu().float(), ref) >= 0.9999
    res = torch.randn(M, N, generator=g)
    gate = torch.randn(N, generator=g)
    x = res.clone().to(dev)
    b200q.gemm_w4a8(qa.to(dev), packed, K, da.to(dev), dw.to(dev), zp.to(dev), rs.to(dev), None,
                    epilogue=b200q.EPI_GATE_RESIDUAL, residual=x, gate=gate.to(dev))
    assert float((x.cpu().double() - (res.double() + y * gate.double())).abs().max()) <= 1e-4


# ---- 2-CTA cluster / TMA-multicast scheduling: identical results ----------------------------------------------
@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (384, 512, 1536), (130, 272, 144), (1000, 1536, 512), (129, 8, 16)])
def test_cluster_modes_bit_exact(dev, M, N, K, mode):
    qa, qw = _codes(M, K, M + K + 5), _codes(N, K, N + K + 6, -128, 127)
    ref = O.int_accumulators(qa, qw)
    try:
        b200q.gemm_set_cluster(mode)
        acc = b200q.gemm_w8a8(qa.to(dev), qw.to(dev), out_dtype=torch.int32)
        assert torch.equal(acc.cpu(), ref)
        # W4 converters + remote barrier arrives
        qw4 = _codes(N, K, 7, -8, 7)
        da = torch.rand(M) * 0.01 + 0.005
        dw = torch.rand(N) * 0.1 + 0.1
        rs = qa.to(torch.int32).sum(dim=1).to(torch.int32)
        out = b200q.gemm_w4a8(qa.to(dev), b200q.pack_w4(qw4.to(dev)), K, da.to(dev), dw.to(dev), None, rs.to(dev), None,
                              out_dtype=torch.float32)
        ref4 = da.double()[:, None] * dw.double()[None, :] * O.int_accumulators(qa, qw4).double()
        assert float(((out.cpu().double() - ref4).abs() / (ref4.abs() + 1.0)).max()) <= 2e-5
        # gate-residual epilogue (3-stage ring, prefetched residual boxes) under clusters
        res = torch.randn(M, N)
        gate = torch.randn(N)
        x = res.clone().to(dev)
        b200q.gemm_w8a8(qa.to(dev), qw.to(dev), da.to(dev), dw.to(dev), None, None, None, epilogue=b200q.EPI_GATE_RESIDUAL,
                        residual=x, gate=gate.to(dev))
        y = da.double()[:, None] * dw.double()[None, :] * ref.double()
        assert float((x.cpu().double() - (res.double() + y * gate.double())).abs().max()) <= 1e-3 * float(y.abs().max() + 1)
    finally:
        b200q.gemm_set_cluster(0)


@pytest.mark.parametrize("M,N,K", [(1500, 512, 1536), (2100, 304, 1008), (1024, 1536, 8960)])
def test_w4a8_expanded_weights_path_equals_in_kernel_converter(dev, M, N, K):
    """M >= 1024 (include/b200q.h expand_ws): the packed weights are expanded once per call and the product runs on the
    W8A8 kernel with the nibble bias folded through the zero point - bit-identical to the in-kernel converter path, for
    every epilogue, incl. a K that is a multiple of 16 but not of 32, and asymmetric weights."""
    g = torch.Generator().manual_seed(M + K)
    qa, qw = _codes(M, K, 31), _codes(N, K, 32, -8, 7)
    da = (torch.rand(M, generator=g) * 0.01 + 0.005).to(dev)
    dw = (torch.rand(N, generator=g) * 0.1 + 0.1).to(dev)
    zp = torch.randint(-3, 4, (N,), generator=g).float().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    rs = qa.to(torch.int32).sum(dim=1).to(torch.int32).to(dev)
    packed = b200q.pack_w4(qw.to(dev))
    res = torch.randn(M, N, generator=g).to(dev)
    gate = torch.randn(N, generator=g).to(dev)
    outs = {}
    for expand in (True, False):
        b200q.w4_expand = expand
        try:
            outs[expand] = (
                b200q.gemm_w4a8(qa.to(dev), packed, K, da, dw, zp, rs, bias, out_dtype=torch.float32),
                b200q.gemm_w4a8(qa.to(dev), packed, K, da, dw, zp, rs, bias, epilogue=b200q.EPI_GELU_TANH),
                b200q.gemm_w4a8(qa.to(dev), packed, K, da, dw, zp, rs, bias, out_dtype=torch.float32,
                                epilogue=b200q.EPI_GATE_RESIDUAL, residual=res.clone(), gate=gate))
        finally:
            b200q.w4_expand = True
    for a, b in zip(outs[True], outs[False]):
        assert torch.equal(a, b)
    acc = O.int_accumulators(qa, qw).double() + zp.cpu().double()[None, :] * rs.cpu().double()[:, None]
    ref = da.cpu().double()[:, None] * dw.cpu().double()[None, :] * acc + bias.cpu().double()[None, :]
    rel = float(((outs[True][0].cpu().double() - ref).abs() / (ref.abs() + 1.0)).max())
    assert rel <= 2e-5, rel
