"""-m gpu: the bf16 tcgen05 flash-attention kernel (include/b200q.h b200q_attn_bf16) against fp32 softmax attention on the
same bf16 inputs = the reference's attention call (wan/modules/attention.py:94-127, SDPA fallback :171-178).
Tolerance: max |err| <= 2e-2 * max|ref| and cosine >= 0.9999 (bf16 P and bf16 output; fp32 accumulation and statistics)."""
import pytest
import torch

import b200q

pytestmark = pytest.mark.gpu


def _ref(q, k, v, H, scale=None):
    Lq, D = q.shape
    hd = D // H
    scale = hd ** -0.5 if scale is None else scale
    qh, kh, vh = (t.float().view(t.shape[0], H, hd).permute(1, 0, 2) for t in (q, k, v))
    s = (qh @ kh.transpose(1, 2)) * scale
    p = torch.softmax(s, dim=-1)
    lse = torch.logsumexp(s, dim=-1) * 1.4426950408889634
    return (p @ vh).permute(1, 0, 2).reshape(Lq, D), lse


def _check(out, ref, tol=2e-2):
    out, ref = out.float(), ref.float()
    err = float((out - ref).abs().max() / ref.abs().max())
    cos = float((out.double().flatten() @ ref.double().flatten()) / (out.double().norm() * ref.double().norm()))
    assert err <= tol and cos >= 0.9999, (err, cos)


@pytest.mark.parametrize("H,Lq,Lk", [(2, 256, 128), (2, 200, 300), (1, 1, 1), (3, 257, 129), (12, 1000, 517), (2, 512, 1024),
                                     (1, 128, 4096), (4, 3000, 512), (2, 77, 2000)])
def test_attn_bf16_matches_fp32_softmax(dev, H, Lq, Lk):
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + Lq + Lk)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g).to(torch.bfloat16) for n in (Lq, Lk, Lk))
    out, lse = b200q.attn_bf16(q, k, v, H, want_lse=True)
    ref, lse_ref = _ref(q, k, v, H)
    _check(out, ref)
    assert torch.allclose(lse, lse_ref, atol=2e-3, rtol=1e-4)
    assert torch.equal(out, b200q.attn_bf16(q, k, v, H, n_splits=1))            # deterministic


def test_attn_bf16_strided_operands_and_scale(dev):
    """q, k, v as column slices of a fused q|k|v GEMM output (row pitch 3*D), a non-default softmax scale"""
    H, L = 3, 700
    D = H * 128
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(L, 3 * D, device=dev, generator=g).to(torch.bfloat16)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    out = b200q.attn_bf16(q, k, v, H, sm_scale=0.05)
    _check(out, _ref(q, k, v, H, 0.05)[0])


def test_attn_bf16_growing_maxima_trigger_rescaling(dev):
    """row maxima that keep growing along the keys (every block exceeds the reference maximum by more than 2^8) and large
    magnitudes: exercises the lazy O / l rescale path in every block, within a block, and the -inf start"""
    H, Lq, Lk = 2, 300, 1500
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.randn(Lq, H * 128, device=dev, generator=g)
    k = torch.randn(Lk, H * 128, device=dev, generator=g)
    v = torch.randn(Lk, H * 128, device=dev, generator=g)
    ramp = torch.linspace(0.2, 6.0, Lk, device=dev).view(-1, 1)
    k = k * ramp                                                     # scores grow with the key index
    k[700:716] *= 3                                                  # a jump inside a 128-key block (chunk-level bump)
    q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    out = b200q.attn_bf16(q, k, v, H)
    _check(out, _ref(q, k, v, H)[0], tol=3e-2)
    # descending maxima: the first block fixes the reference, nothing rescales afterwards
    out = b200q.attn_bf16(q, k.flip(0).contiguous(), v, H)
    _check(out, _ref(q, k.flip(0), v, H)[0], tol=3e-2)


def test_attn_bf16_many_items_persistent_schedule(dev):
    """more (head, query-tile) items than SMs: every CTA walks several items, barriers wrap their phases many times"""
    H, Lq, Lk = 12, 4096 + 100, 640
    g = torch.Generator(device="cuda").manual_seed(3)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g).to(torch.bfloat16) for n in (Lq, Lk, Lk))
    _check(b200q.attn_bf16(q, k, v, H), _ref(q, k, v, H)[0])


def test_attn_bf16_rejects_bad_arguments(dev):
    q = torch.randn(64, 256, device=dev).to(torch.bfloat16)
    with pytest.raises(b200q.B200QError):
        b200q.attn_bf16(q, q, q, 4)                                  # head_dim 64
    with pytest.raises(b200q.B200QError):
        b200q.attn_bf16(q.float(), q, q, 2)                          # not bf16
    with pytest.raises(b200q.B200QError):
        b200q.attn_bf16(q, q[:, :128], q, 2)                         # shape mismatch


def test_block_with_own_attention_core_matches_library_core(dev):
    """the DiT block with ATTENTION_CORE = b200q vs library SDPA: same function, bf16-level agreement"""
    from wan_b200 import model as M
    cfg = M.WanConfig(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=64)
    dit = M.WanDiTQ.random(cfg, seed=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    lat = torch.randn(16, 3, 8, 12, device=dev, generator=g)
    ctx = torch.randn(20, 64, device=dev, generator=g)
    t = torch.tensor([500.0], device=dev)
    try:
        M.set_attention_core("library")
        y_lib = dit.forward(lat, t, ctx)
        M.set_attention_core("b200q")
        y_own = dit.forward(lat, t, ctx)
    finally:
        M.set_attention_core("library")
    _check(y_own, y_lib, tol=3e-2)


@pytest.mark.parametrize("H,Lq,Lk,S", [(2, 300, 2500, 2), (3, 700, 4096 + 77, 3), (1, 256, 1024, 4), (2, 513, 1100, 8)])
def test_attn_bf16_key_splits_merge(dev, H, Lq, Lk, S):
    """work items split along the keys + log-sum-exp merge = the unsplit result up to bf16 rounding of the partials"""
    g = torch.Generator(device="cuda").manual_seed(S * 100 + Lq)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g).to(torch.bfloat16) for n in (Lq, Lk, Lk))
    out = b200q.attn_bf16(q, k, v, H, n_splits=S)
    _check(out, _ref(q, k, v, H)[0])
    one = b200q.attn_bf16(q, k, v, H, n_splits=1)
    assert float((out.float() - one.float()).abs().max()) <= 2e-2 * float(one.float().abs().max())
    # the library's own proposal is a valid count and small problems are left alone
    assert b200q.load().b200q_attn_bf16_splits(Lq, Lk, H) in (1, 2, 3, 4)
    assert b200q.load().b200q_attn_bf16_splits(64, 512, 2) == 1


def test_attn_bf16_bounded_and_unbounded_heads_in_one_call(dev):
    """head classes (include/b200q.h qk_norm_ws): head 0 has small norms (Cauchy-Schwarz bound <= 80 -> max-free kernel),
    head 1 has keys scaled far beyond it (online-softmax kernel), head 2 sits just above the bound; one call, every head
    must match fp32 softmax attention, with and without key splits"""
    H, Lq, Lk = 3, 700, 1300
    g = torch.Generator(device="cuda").manual_seed(21)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g) for n in (Lq, Lk, Lk))
    k[:, 128:256] *= 12.0
    k[:, 256:] *= 3.5
    q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    ref, lse_ref = _ref(q, k, v, H)
    out, lse = b200q.attn_bf16(q, k, v, H, want_lse=True)
    _check(out, ref, tol=3e-2)
    assert torch.allclose(lse, lse_ref, atol=5e-3, rtol=1e-4)
    _check(b200q.attn_bf16(q, k, v, H, n_splits=3), ref, tol=3e-2)


@pytest.mark.parametrize("poly", [2, 3, 4, 5])
def test_attn_bf16_max_free_kernel_equals_online_softmax(dev, poly):
    """the max-free kernel (bounded heads) and the online-softmax kernel compute the same function: compare them on the
    same inputs for every polynomial share, full and partial last key block, rows beyond Lq"""
    H, Lq, Lk = 4, 1000, 2100
    g = torch.Generator(device="cuda").manual_seed(poly)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g).to(torch.bfloat16) for n in (Lq, Lk, Lk))
    try:
        b200q.attn_bf16_set_fast(-1)
        online, lse_o = b200q.attn_bf16(q, k, v, H, want_lse=True)
        b200q.attn_bf16_set_fast(poly)
        fast, lse_f = b200q.attn_bf16(q, k, v, H, want_lse=True)
    finally:
        b200q.attn_bf16_set_fast(-2)
    ref, lse_ref = _ref(q, k, v, H)
    _check(fast, ref)
    _check(online, ref)
    assert float((fast.float() - online.float()).abs().max()) <= 1e-2 * float(ref.abs().max())
    assert torch.allclose(lse_f, lse_ref, atol=2e-3, rtol=1e-4) and torch.allclose(lse_o, lse_ref, atol=2e-3, rtol=1e-4)


def test_attn_bf16_nan_and_huge_inputs_take_the_online_kernel(dev):
    """a head with an enormous key norm must not overflow in the max-free kernel: it is classified unbounded"""
    H, Lq, Lk = 2, 256, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g) for n in (Lq, Lk, Lk))
    k[100, :128] *= 300.0                                           # one giant key in head 0
    q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    out = b200q.attn_bf16(q, k, v, H)
    assert torch.isfinite(out.float()).all()
    _check(out, _ref(q, k, v, H)[0], tol=3e-2)


@pytest.mark.parametrize("variant,cl", [(0, 1), (0, 2), (1, 2)])
@pytest.mark.parametrize("H,Lq,Lk", [(1, 1, 1), (2, 200, 300), (3, 513, 129), (12, 1000, 517), (1, 128, 4096), (2, 77, 2000),
                                     (3, 5000, 3000)])
def test_attn_bf16_kernel_variants(dev, variant, cl, H, Lq, Lk):
    """the three kernels behind b200q_attn_bf16 - two-tile on single CTAs, two-tile on CTA pairs (cta_group::2), Q-resident
    on CTA pairs - compute the same function: each against fp32 softmax attention (output and log-sum-exp), with and
    without key splits, partial last key block, rows beyond Lq, one key block, odd block counts"""
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + Lq + Lk)
    q, k, v = (torch.randn(n, H * 128, device=dev, generator=g).to(torch.bfloat16) for n in (Lq, Lk, Lk))
    ref, lse_ref = _ref(q, k, v, H)
    try:
        b200q.attn_bf16_set_cluster(cl)
        b200q.attn_bf16_set_variant(variant)
        out, lse = b200q.attn_bf16(q, k, v, H, want_lse=True)
        again = b200q.attn_bf16(q, k, v, H, n_splits=1)
        split = b200q.attn_bf16(q, k, v, H, n_splits=2) if Lk >= 2000 else None
        b200q.attn_bf16_set_fast(-1)                                # every head through the online-softmax kernel
        online = b200q.attn_bf16(q, k, v, H)
    finally:
        b200q.attn_bf16_set_fast(-2)
        b200q.attn_bf16_set_cluster(2)
        b200q.attn_bf16_set_variant(1)
    _check(out, ref)
    _check(online, ref)
    assert torch.allclose(lse, lse_ref, atol=2e-3, rtol=1e-4)
    assert torch.equal(out, again)
    if split is not None:
        _check(split, ref)


def test_rmsnorm_rope_head_norm_maxima_feed_the_attention_classification(dev):
    """b200q_rmsnorm_rope_stats leaves max_i |q_i,h|^2 of its bf16 output per 128-wide head (accumulating over calls);
    b200q_attn_bf16_prenorm classifies the heads from those instead of re-reading q and k: same kernels, same result.
    A head pushed over the bound (huge key rows) must take the online-softmax kernel through this path too."""
    H, L, Lk = 6, 777, 1300
    D = H * 128
    g = torch.Generator(device="cuda").manual_seed(11)
    xq = torch.randn(L, 3 * D, device=dev, generator=g).to(torch.bfloat16)[:, :D]          # strided source, as in the block
    xk = torch.randn(Lk, D, device=dev, generator=g).to(torch.bfloat16)
    xk[5, 256:384] *= 60.0                                                                  # head 2: one dominant key row
    v = torch.randn(Lk, D, device=dev, generator=g).to(torch.bfloat16)
    wq, wk = torch.rand(D, device=dev, generator=g) + 0.5, (torch.rand(D, device=dev, generator=g) + 0.5) * 3.0
    cos = torch.rand(L, 64, device=dev, generator=g); sin = torch.rand(L, 64, device=dev, generator=g)
    nrm = torch.zeros(2 * H, device=dev)
    q = b200q.rmsnorm_rope(xq, wq, 1e-6, cos, sin, 128, head_sq_max=nrm[:H])
    k = b200q.rmsnorm_rope(xk, wk, 1e-6, None, None, 0, head_sq_max=nrm[H:])
    assert torch.equal(q, b200q.rmsnorm_rope(xq, wq, 1e-6, cos, sin, 128))                  # the output itself is unchanged
    want = torch.cat([q.float().view(L, H, 128).square().sum(-1).amax(0), k.float().view(Lk, H, 128).square().sum(-1).amax(0)])
    assert torch.allclose(nrm, want, rtol=1e-5, atol=0)
    half = nrm.clone()
    b200q.rmsnorm_rope(xq[: L // 2], wq, 1e-6, cos[: L // 2], sin[: L // 2], 128, head_sq_max=half[:H])   # accumulates: max stays
    assert torch.equal(half, nrm)
    out_pre, lse_pre = b200q.attn_bf16(q, k, v, H, want_lse=True, qk_sq_max=nrm)
    out, lse = b200q.attn_bf16(q, k, v, H, want_lse=True)
    assert torch.equal(out_pre, out) and torch.equal(lse_pre, lse)
    bound = (nrm[:H] * nrm[H:]).sqrt() * 128 ** -0.5 * 1.4426950408889634
    assert (bound > 80).any() and (bound <= 80).any()                                        # both kernels were exercised
    _check(out_pre, _ref(q, k, v, H)[0], tol=3e-2)
    with pytest.raises(b200q.B200QError):
        b200q.attn_bf16(q, k, v, H, qk_sq_max=nrm[:H])
    with pytest.raises(b200q.B200QError):
        b200q.rmsnorm_rope(xq, wq, 1e-6, cos, sin, 128, head_sq_max=nrm)
