"""Test-only stand-in for the b200q kernel ops, backed by the CPU oracle, so the HOST logic of the qdiff
mirror (module surgery, regex selection, parameter dicts, bit-width refactor) can be exercised in the
`-m "not gpu"` suite.  The product has no CPU path: without this monkeypatch the same calls raise."""
import torch

from oracle import fakequant_oracle as O


def quant_rows(x, n_bits=8, sym=True, dynamic=True, want_rowsum=True, out=None, want_stats=False):
    q, d, z = O.quant_rows(x.float(), n_bits, sym, dynamic)
    qi = q.clamp(-128, 127).to(torch.int8)
    rs = qi.to(torch.int32).sum(dim=1).to(torch.int32) if want_rowsum else None
    res = (qi, d.flatten(), z.flatten(), rs)
    if want_stats:
        xf = x.float()
        if sym:
            return res + (xf.abs().max(dim=1)[0], None)
        return res + (xf.max(dim=1)[0].clamp_min(0), xf.min(dim=1)[0].clamp_max(0))
    return res


def quant_rows_static(x, delta, zero_point, n_bits=8, sym=False, want_rowsum=False):
    q = O.quant_codes_rows(x.float(), delta.reshape(-1, 1).float(), zero_point.reshape(-1, 1).float(), n_bits, sym)
    qi = q.clamp(-128, 127).to(torch.int8)
    return qi, (qi.to(torch.int32).sum(dim=1).to(torch.int32) if want_rowsum else None)


def dequant_rows(q, delta, zero_point=None, out_dtype=torch.float32):
    z = 0 if zero_point is None else zero_point.reshape(-1, 1).float()
    return ((q.float() + z) * delta.reshape(-1, 1).float()).to(out_dtype)


def gemm_w8a8(qa, qw, delta_a=None, delta_w=None, zp_w=None, rowsum_a=None, bias=None, out_dtype=torch.bfloat16,
              epilogue=0, residual=None, gate=None, out=None):
    acc = O.int_accumulators(qa, qw)
    if out_dtype == torch.int32:
        return acc
    full = acc.double()
    if zp_w is not None:
        full = full + zp_w.double()[None, :] * rowsum_a.double()[:, None]
    y = delta_a.double()[:, None] * delta_w.double()[None, :] * full
    if bias is not None:
        y = y + bias.double()
    return y.to(out_dtype)


def pack_w4(codes):
    return codes.clone()          # opaque handle for the fake gemm_w4a8


def gemm_w4a8(qa, qw4, K, *args, **kwargs):
    return gemm_w8a8(qa, qw4, *args, **kwargs)


def had_transform(x, colscale=None, hadK=None, K=1, log2_width=0):
    """float64 restatement of what b200q_had_quant_rows applies in front of its quantizer (include/b200q.h)."""
    y = x.double()
    if colscale is not None:
        y = y * colscale.double().reshape(1, -1)
    if log2_width > 0:
        W = 1 << log2_width
        y = y.reshape(-1, K, W)
        h = 1
        while h < W:
            y = y.reshape(-1, K, W // (2 * h), 2, h)
            y = torch.stack((y[..., 0, :] + y[..., 1, :], y[..., 0, :] - y[..., 1, :]), dim=-2)
            h *= 2
        y = y.reshape(-1, K, W)
        if K > 1:
            y = torch.einsum("ij,bjk->bik", hadK.double().reshape(K, K), y)
        y = y.reshape(x.shape)
    return y


def had_quant_rows(x, colscale=None, hadK=None, K=1, log2_width=0, n_bits=8, want_rowsum=True, want_y=False, out=None):
    y = had_transform(x, colscale, hadK, K, log2_width).float()
    q, d, _, rs = quant_rows(y, n_bits, True, True, want_rowsum)
    return q, d, rs, (y if want_y else None)


def install(monkeypatch):
    import b200q
    import qdiff.base.base_quantizer as bq
    for name in ("quant_rows", "quant_rows_static", "dequant_rows", "gemm_w8a8", "pack_w4", "gemm_w4a8", "had_quant_rows"):
        monkeypatch.setattr(b200q, name, globals()[name])
    monkeypatch.setattr(bq, "_on_cuda", lambda x: (x, None))
