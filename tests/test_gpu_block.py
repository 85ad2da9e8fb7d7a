"""-m gpu: one quantized Wan DiT block through the integer runtime vs the fp32 fake-quant oracle block
(oracle.fakequant_oracle.WanBlockOracle = restated wan/modules/model.py:293-370 with the imported-reference-pinned
quantizers).  Tolerance (north_star): cosine >= 0.999 on the block output; max relative error stated below."""
import pytest
import torch

import b200q
from oracle import fakequant_oracle as O
from wan_b200 import model as M

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()))


@pytest.mark.parametrize("dim,ffn,heads,grid", [(256, 512, 2, (2, 6, 8)), (384, 1024, 3, (3, 4, 4))])
def test_block_matches_oracle(dev, dim, ffn, heads, grid):
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=heads, num_layers=1)
    p = O.make_block_params(dim, ffn, seed=0)
    L, T = grid[0] * grid[1] * grid[2], 32
    g = torch.Generator().manual_seed(1)
    x = torch.randn(L, dim, generator=g)
    e = torch.randn(6, dim, generator=g) * 0.1
    ctx = torch.randn(T, dim, generator=g)
    ref = O.WanBlockOracle(p, dim, ffn, heads).forward(x.clone(), e, grid, ctx)

    blk = M.WanBlockQ.from_fp_params(cfg, p)
    cos, sin = M.rope_table(dim // heads, grid, dev)
    out = blk.forward(x.clone().to(dev), e.to(dev), ctx.to(dev), cos, sin).cpu()
    c = _cos(out, ref)
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert c >= 0.999, c
    assert rel <= 5e-2, rel          # bf16 intermediates (q,k,v, attention, GELU output) vs the oracle's fp32


def test_block_weight_codes_bit_exact(dev):
    dim, ffn = 256, 512
    p = O.make_block_params(dim, ffn, seed=3)
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=2, num_layers=1)
    blk = M.WanBlockQ.from_fp_params(cfg, p)
    for name, w in (("ffn.0", blk.w_f0), ("ffn.2", blk.w_f2), ("self_attn.o", blk.w_o)):
        q, d, z = O.quant_rows(p[name + ".weight"], 8, False, dynamic=False)
        assert torch.equal(w.codes.cpu().float(), q) and torch.equal(w.delta.cpu(), d.flatten())
        assert torch.equal(w.zp.cpu(), z.flatten())
    q, d, z = O.quant_rows(p["self_attn.k.weight"], 8, False, dynamic=False)
    assert torch.equal(blk.w_qkv.codes[dim:2 * dim].cpu().float(), q)


def test_dit_forward_runs_and_is_deterministic(dev):
    cfg = M.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=64)
    dit = M.WanDiTQ.random(cfg, seed=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randn(16, 3, 8, 12, device=dev, generator=g)
    ctx = torch.randn(20, 64, device=dev, generator=g)
    t = torch.tensor([500.0], device=dev)
    y1 = dit.forward(lat, t, ctx)
    y2 = dit.forward(lat, t, ctx)
    assert y1.shape == lat.shape and torch.isfinite(y1).all()
    assert torch.equal(y1, y2)


def test_calibration_collector(dev):
    import torch.nn as nn
    from wan_b200.calibration import CalibrationCollector
    m = nn.Sequential(nn.Linear(64, 128), nn.GELU(), nn.Linear(128, 32)).to(dev)
    col = CalibrationCollector(m)
    xs = [torch.randn(3, 50, 64, device=dev) * (i + 1) for i in range(3)]
    ref0, ref1 = [], []
    for x in xs:
        ref0.append(O.calib_absmax(x.cpu()))
        ref1.append(O.calib_absmax(m[1](m[0](x)).detach().cpu()))
        m(x)
    sd = col.state_dict()
    assert torch.equal(sd["0"].max(dim=0)[0], O.calib_merge(ref0))
    assert torch.equal(sd["2"].max(dim=0)[0], O.calib_merge(ref1))


@pytest.mark.parametrize("L,heads,hd,grid", [(96, 2, 128, (2, 6, 8)), (60, 3, 64, (3, 4, 5))])
def test_rmsnorm_rope_kernel(dev, L, heads, hd, grid):
    """Fused RMSNorm(full dim)+RoPE vs the oracle restatement of model.py:43-89 (float64 complex RoPE)."""
    D = heads * hd
    g = torch.Generator().manual_seed(L)
    x = (torch.randn(L, 3 * D, generator=g) * 2).to(torch.bfloat16)
    w = torch.rand(D, generator=g) + 0.5
    xs = x[:, D:2 * D]                                             # strided slice, like k inside the fused qkv output
    cos, sin = M.rope_table(hd, grid, dev)
    out = b200q.rmsnorm_rope(xs.to(dev) if False else x.to(dev)[:, D:2 * D], w.to(dev), 1e-6, cos, sin, hd).cpu().float()
    ref = O.rms_norm(xs, w, 1e-6)                                  # bf16 in -> fp32 (type_as(x) * fp32 weight)
    ref = O.rope_apply(ref.view(L, heads, hd), grid, O.wan_freqs(hd)).reshape(L, D)
    assert float((out - ref).abs().max()) <= 2e-2 * float(ref.abs().max())     # bf16 output rounding
    c = float((out.double().flatten() @ ref.double().flatten()) / (out.double().norm() * ref.double().norm()))
    assert c >= 0.99999
    out2 = b200q.rmsnorm_rope(x.to(dev)[:, :D], w.to(dev), 1e-6).cpu().float()
    assert float((out2 - O.rms_norm(x[:, :D], w, 1e-6)).abs().max()) <= 2e-2 * float(out2.abs().max())


def test_quantized_attention_parity_mode(dev, golden_dir):
    """Materialised quantized attention (quant_opensora.py:430-478, 'row' attn-map group) vs the golden vectors of the
    imported reference quantizers: Q/K/V deltas and dequantised tensors bit-exact; P codes within one step (S comes
    from the int8 GEMM's fp32 epilogue instead of an fp32 matmul); output cosine >= 0.9999."""
    import os
    from wan_b200.attention_q import quantized_attention_parity
    rec = torch.load(os.path.join(golden_dir, "quant_attention.pt"))
    B, H, L, hd = rec["q"].shape
    to_tok = lambda t: t[0].permute(1, 0, 2).reshape(L, H * hd).contiguous().to(dev)      # [1,H,L,hd] -> [L, H*hd]
    out, info = quantized_attention_parity(to_tok(rec["q"]), to_tok(rec["k"]), to_tok(rec["v"]), H, return_info=True)
    # reference deltas are [B*H*L, 1] in (h, l) order; ours are [L, H]
    assert torch.equal(info["dq"].t().reshape(-1).cpu(), rec["q_delta"].flatten())
    assert torch.equal(info["dk"].t().reshape(-1).cpu(), rec["k_delta"].flatten())
    assert torch.equal(info["dv"].cpu(), rec["v_delta"].flatten())
    q_dq = (info["qq"].float() * info["dq"].repeat_interleave(hd, dim=1)).cpu()
    assert torch.equal(q_dq, to_tok(rec["q_dequant"]).cpu())
    for h in range(H):
        pdq = ((info["pq"][h].float() + info["zp"][h][:, None]) * info["dp"][h][:, None]).t().cpu()   # [Lq, Lk]
        ref = rec["attn_quant"][0, h]
        step = rec["p_delta"].reshape(H, L)[h]
        assert float(((pdq - ref).abs() / step[None, :]).max()) <= 1.0 + 1e-3
    ref_out = rec["out"][0].permute(1, 0, 2).reshape(L, H * hd)
    c = float((out.cpu().double().flatten() @ ref_out.double().flatten()) / (out.cpu().double().norm() * ref_out.double().norm()))
    assert c >= 0.9999, c


@pytest.mark.parametrize("dim,ffn,heads,grid", [(256, 512, 2, (2, 6, 8)), (384, 1024, 3, (3, 7, 13))])
def test_block_int8_attention_matches_oracle(dev, dim, ffn, heads, grid):
    """W8A8 linears + 8-bit Q.K^T / P.V attention (BASELINE configs[4] semantics) through the fused tcgen05 attention
    kernel vs the oracle block whose attention is the row-step fake-quant attention (oracle pinned to the imported
    reference quantizers, tests/golden/quant_attention_rowstep.pt)."""
    cfg = M.WanConfig(dim=dim, ffn_dim=ffn, num_heads=heads, num_layers=1)
    p = O.make_block_params(dim, ffn, seed=0)
    L, T = grid[0] * grid[1] * grid[2], 40
    g = torch.Generator().manual_seed(1)
    x = torch.randn(L, dim, generator=g)
    e = torch.randn(6, dim, generator=g) * 0.1
    ctx = torch.randn(T, dim, generator=g)
    ref = O.WanBlockOracle(p, dim, ffn, heads, attn_quant=dict(mode="rowstep")).forward(x.clone(), e, grid, ctx)
    blk = M.WanBlockQ.from_fp_params(cfg, p, attn_quant=True)
    cos, sin = M.rope_table(dim // heads, grid, dev)
    n0 = b200q.launch_count
    out = blk.forward(x.clone().to(dev), e.to(dev), ctx.to(dev), cos, sin).cpu()
    assert b200q.launch_count - n0 >= 20
    c = _cos(out, ref)
    rel = float((out - ref).abs().max() / ref.abs().max())
    assert c >= 0.999, c
    assert rel <= 5e-2, rel


@pytest.mark.parametrize("L,heads,grid", [(96, 2, (2, 6, 8)), (273, 3, (3, 7, 13))])
def test_rmsnorm_rope_quant_kernel(dev, L, heads, grid):
    """RMSNorm+RoPE with the fused per-(token, head) quantizer vs the oracle (float64 RoPE, then DynamicQuantizer on the
    [L*H, 128] view): scales to 1e-5 relative, codes within one step (the kernel rotates in fp32), mismatch rate < 1 %;
    the optional bf16 output is the plain rmsnorm_rope output."""
    hd = 128
    D = heads * hd
    g = torch.Generator().manual_seed(L)
    x = (torch.randn(L, 3 * D, generator=g) * 2).to(torch.bfloat16)
    w = torch.rand(D, generator=g) + 0.5
    cos, sin = M.rope_table(hd, grid, dev)
    xs = x.to(dev)[:, D:2 * D]
    q, dq, y = b200q.rmsnorm_rope_quant(xs, w.to(dev), 1e-6, cos, sin, hd, want_bf16=True)
    ref = O.rope_apply(O.rms_norm(x[:, D:2 * D], w, 1e-6).view(L, heads, hd), grid, O.wan_freqs(hd))     # [L, H, hd] fp32
    qo, do, _ = O.quant_rows(ref.reshape(L * heads, hd), 8, True, True)
    assert torch.allclose(dq.cpu().flatten(), do.flatten(), rtol=1e-5)
    diff = (q.cpu().float().view(L * heads, hd) - qo).abs()
    assert diff.max() <= 1 and float((diff > 0).float().mean()) < 0.01
    assert torch.equal(y.cpu(), b200q.rmsnorm_rope(xs, w.to(dev), 1e-6, cos, sin, hd).cpu())
    # no RoPE (cross-attention q/k): codes of the RMSNorm output
    q2, dq2, _ = b200q.rmsnorm_rope_quant(xs, w.to(dev), 1e-6, None, None, hd)
    qo2, do2, _ = O.quant_rows(O.rms_norm(x[:, D:2 * D], w, 1e-6).reshape(L * heads, hd), 8, True, True)
    assert torch.allclose(dq2.cpu().flatten(), do2.flatten(), rtol=1e-6)
    d2 = (q2.cpu().float().view(L * heads, hd) - qo2).abs()
    assert d2.max() <= 1 and float((d2 > 0).float().mean()) < 0.002
