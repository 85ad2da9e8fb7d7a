"""b200q — ctypes binding of libb200q.so (C ABI in include/b200q.h).

PyTorch here is plumbing only: it owns device memory and streams; every op below hands raw
`data_ptr()`s, sizes and the current CUDA stream to the hand-written sm_100a kernels.
There is NO CPU / eager fallback: a missing library or a non-CUDA tensor raises.

Mirrors what the reference reaches through `viditq_extension.{fused,qgemm}`
(ViDiT-Q/kernels/csrc/fused/pybind.cpp:56-99, kernels/csrc/qgemm/pybind.cpp:5-12).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libb200q.so")

F32, BF16, F16, I32 = 0, 1, 2, 3
EPI_NONE, EPI_GELU_TANH, EPI_GATE_RESIDUAL = 0, 1, 2

_DTYPE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16, torch.int32: I32}

_lib = None
# incremented on every successful kernel-launching call (bench.py reports it as gpu_launches)
launch_count = 0


class B200QError(RuntimeError):
    pass


_SIGNATURES = {
    "b200q_version": (c_int, []),
    "b200q_last_error": (c_char_p, []),
    "b200q_device_info": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b200q_quant_rows": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                 c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200q_quant_rows_static": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "b200q_dequant_rows": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int64, c_void_p]),
    "b200q_calib_absmax_minmax": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64,
                                          c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200q_gemm_w8a8": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "b200q_gemm_w4a8": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200q_gemm_set_cluster": (c_int, [c_int]),
    "b200q_pack_w4": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "b200q_ln_mod_quant": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_float,
                                   c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int64, c_void_p]),
    "b200q_rmsnorm_rope": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_float, c_void_p, c_void_p,
                                   c_int, c_void_p, c_int64, c_void_p]),
    "b200q_quant_vt": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                               c_void_p]),
    "b200q_attn_i8": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                              c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_void_p, c_int,
                              c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "b200q_attn_set_mode": (c_int, [c_int]),
    "b200q_scatter_rows": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "b200q_attn_bf16_set_mode": (c_int, [c_int]),
    "b200q_attn_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float,
                                c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200q_attn_bf16_prenorm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_float,
                                        c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200q_rmsnorm_rope_stats": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                         c_void_p, c_int64, c_void_p, c_void_p]),
    "b200q_attn_bf16_set_fast": (c_int, [c_int]),
    "b200q_attn_bf16_set_cluster": (c_int, [c_int]),
    "b200q_attn_bf16_set_variant": (c_int, [c_int]),
    "b200q_attn_bf16_splits": (c_int, [c_int64, c_int64, c_int]),
    "b200q_rmsnorm_rope_quant": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_float, c_void_p, c_void_p,
                                         c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "b200q_had_quant_rows": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b200q_had_set_mode": (c_int, [c_int]),
    "b200q_gate_residual": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                    c_int64, c_int64, c_void_p]),
}


def exported_symbols():
    """Names every build of the library must export (tests compare with include/b200q.h)."""
    return sorted(_SIGNATURES)


def load():
    """dlopen the library once; fail loudly when it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200QError(
            f"{LIB_PATH} not found: build it with `python wan2.1-quantization_b200/build.py` "
            "(there is no CPU fallback for the quantized hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the build lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def version():
    v = load().b200q_version()
    return (v >> 16, (v >> 8) & 0xFF, v & 0xFF)


def _check(rc, what):
    global launch_count
    if rc != 0:
        msg = load().b200q_last_error().decode("utf-8", "replace")
        raise B200QError(f"{what} failed (status {rc}): {msg}")
    launch_count += 1


def _cuda(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise B200QError(f"{name}: expected a CUDA tensor — libb200q has no CPU path")
    return t


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _rows2d(x, name):
    _cuda(x, name)
    if x.dim() != 2:
        raise B200QError(f"{name}: expected a 2-D [rows, cols] tensor, got {tuple(x.shape)}")
    if x.stride(1) != 1 and x.shape[1] > 1:
        x = x.contiguous()
    return x


def _ld(x):
    """leading dimension (elements) of a 2-D row-major view"""
    return x.stride(0) if x.shape[0] > 1 else max(int(x.shape[1]), 1)


# ---------------------------------------------------------------------------------------------
# (a) per-row quantizer
# ---------------------------------------------------------------------------------------------
def quant_rows(x, n_bits=8, sym=True, dynamic=True, want_rowsum=True, out=None, want_stats=False):
    """x [rows, cols] (fp32|bf16|fp16) -> (codes int8 [rows, cols], delta f32 [rows], zero_point f32 [rows],
    rowsum int32 [rows] | None).  base_quantizer.py:58-99 (dynamic=False) / :110-157 (dynamic=True).
    want_stats=True appends (stat_max, stat_min): sym -> (absmax, None); asym -> (max(rowmax,0), min(rowmin,0))."""
    x = _rows2d(x, "quant_rows")
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise B200QError(f"quant_rows: unsupported dtype {x.dtype}")
    rows, cols = x.shape
    q = out if out is not None else torch.empty((rows, cols), dtype=torch.int8, device=x.device)
    delta = torch.empty(rows, dtype=torch.float32, device=x.device)
    zp = torch.empty(rows, dtype=torch.float32, device=x.device)
    rs = torch.empty(rows, dtype=torch.int32, device=x.device) if want_rowsum else None
    smax = torch.empty(rows, dtype=torch.float32, device=x.device) if want_stats else None
    smin = torch.empty(rows, dtype=torch.float32, device=x.device) if (want_stats and not sym) else None
    rc = load().b200q_quant_rows(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), int(n_bits), int(bool(sym)),
                                 int(bool(dynamic)), _ptr(q), _ld(q), _ptr(delta), _ptr(zp), _ptr(rs),
                                 _ptr(smax), _ptr(smin), _stream())
    _check(rc, "b200q_quant_rows")
    if want_stats:
        return q, delta, zp, rs, smax, smin
    return q, delta, zp, rs


def quant_rows_static(x, delta, zero_point, n_bits=8, sym=False, want_rowsum=False):
    """Quantize with given per-row parameters (base_quantizer.py:63-68)."""
    x = _rows2d(x, "quant_rows_static")
    rows, cols = x.shape
    delta = _cuda(delta, "delta").reshape(-1).float().contiguous()
    zero_point = _cuda(zero_point, "zero_point").reshape(-1).float().contiguous()
    if delta.numel() != rows or zero_point.numel() != rows:
        raise B200QError("quant_rows_static: delta / zero_point must have one entry per row")
    q = torch.empty((rows, cols), dtype=torch.int8, device=x.device)
    rs = torch.empty(rows, dtype=torch.int32, device=x.device) if want_rowsum else None
    rc = load().b200q_quant_rows_static(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), int(n_bits), int(bool(sym)),
                                        _ptr(delta), _ptr(zero_point), _ptr(q), _ld(q), _ptr(rs), _stream())
    _check(rc, "b200q_quant_rows_static")
    return q, rs


def dequant_rows(q, delta, zero_point=None, out_dtype=torch.float32):
    """(q + zero_point) * delta  (base_quantizer.py:159-162)."""
    q = _rows2d(q, "dequant_rows")
    rows, cols = q.shape
    delta = _cuda(delta, "delta").reshape(-1).float().contiguous()
    zp = None if zero_point is None else _cuda(zero_point, "zero_point").reshape(-1).float().contiguous()
    out = torch.empty((rows, cols), dtype=out_dtype, device=q.device)
    rc = load().b200q_dequant_rows(_ptr(q), _ld(q), rows, cols, _ptr(delta), _ptr(zp), _ptr(out), _DTYPE[out_dtype],
                                   _ld(out), _stream())
    _check(rc, "b200q_dequant_rows")
    return out


def had_quant_rows(x, colscale=None, hadK=None, K=1, log2_width=0, n_bits=8, want_rowsum=True, want_y=False, out=None):
    """Fused smooth-scale + structured Hadamard rotation + per-token symmetric quantizer (include/b200q.h, f-2):
    y = (x*colscale) . (H_K (x) H_{2^log2_width}) -> (codes int8, delta f32 [rows], rowsum int32 | None, y f32 | None)."""
    x = _rows2d(x, "had_quant_rows")
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise B200QError(f"had_quant_rows: unsupported dtype {x.dtype}")
    rows, cols = x.shape
    dev = x.device
    q = out if out is not None else torch.empty((rows, cols), dtype=torch.int8, device=dev)
    delta = torch.empty(rows, dtype=torch.float32, device=dev)
    rs = torch.empty(rows, dtype=torch.int32, device=dev) if want_rowsum else None
    y = torch.empty((rows, cols), dtype=torch.float32, device=dev) if want_y else None
    for t, n, want in ((colscale, "colscale", cols), (hadK, "hadK", K * K)):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != want):
            raise B200QError(f"had_quant_rows: {n} must be a contiguous fp32 CUDA tensor of {want} elements")
    rc = load().b200q_had_quant_rows(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), _ptr(colscale), _ptr(hadK), int(K),
                                     int(log2_width), int(n_bits), _ptr(q), _ld(q), _ptr(delta), _ptr(rs), _ptr(y),
                                     _ld(y) if y is not None else 0, _stream())
    _check(rc, "b200q_had_quant_rows")
    return q, delta, rs, y


# ---------------------------------------------------------------------------------------------
# (d) calibration reduction
# ---------------------------------------------------------------------------------------------
def calib_update(x, absmax=None, xmin=None, xmax=None):
    """Running per-channel |x| max / min / max over the rows of x [rows, cols], in place
    (get_calib_data_wanx.py:262-263 + the merge :443-468 / ptq_wanx.py:336)."""
    x = _rows2d(x, "calib_update")
    rows, cols = x.shape
    for t in (absmax, xmin, xmax):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or t.numel() != cols or not t.is_contiguous()):
            raise B200QError("calib_update: statistics buffers must be contiguous fp32 CUDA tensors of [cols]")
    rc = load().b200q_calib_absmax_minmax(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), _ptr(absmax), _ptr(xmin),
                                          _ptr(xmax), _stream())
    _check(rc, "b200q_calib_absmax_minmax")
    return absmax, xmin, xmax


# ---------------------------------------------------------------------------------------------
# (b) quantized linear
# ---------------------------------------------------------------------------------------------
_w4_expand_ws = {}          # device -> reusable scratch for gemm_w4a8's once-per-call weight expansion (stream-ordered reuse)


def _w4_scratch(device, nbytes):
    ws = _w4_expand_ws.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _w4_expand_ws[device] = ws
    return ws


def _gemm(fn_name, qa, qw, N, K, delta_a, delta_w, zp_w, rowsum_a, bias, out_dtype, epilogue, residual, gate, out, extra=()):
    _cuda(qa, "qa"); _cuda(qw, "qw")
    M = qa.shape[0]
    if out is None:
        if epilogue == EPI_GATE_RESIDUAL:
            out = residual
        else:
            out = torch.empty((M, N), dtype=out_dtype, device=qa.device)
    bias_dt = _DTYPE[bias.dtype] if bias is not None else F32
    fn = getattr(load(), fn_name)
    rc = fn(_ptr(qa), _ld(qa), _ptr(qw), _ld(qw), _ptr(delta_a), _ptr(delta_w), _ptr(zp_w), _ptr(rowsum_a),
            _ptr(bias), bias_dt, _ptr(out), _DTYPE[out.dtype], _ld(out), M, N, K, int(epilogue),
            _ptr(residual), _ld(residual) if residual is not None else 0, _ptr(gate), *extra, _stream())
    _check(rc, fn_name)
    return out


def gemm_w8a8(qa, qw, delta_a=None, delta_w=None, zp_w=None, rowsum_a=None, bias=None,
              out_dtype=torch.bfloat16, epilogue=EPI_NONE, residual=None, gate=None, out=None):
    """out[m,n] = epi(da[m]*dw[n]*(sum_k qa[m,k]*qw[n,k] + zp_w[n]*rowsum_a[m]) + bias[n]).
    qa int8 [M,K], qw int8 [N,K].  out_dtype torch.int32 -> raw accumulators."""
    K = qa.shape[1]
    if qw.shape[1] != K:
        raise B200QError(f"gemm_w8a8: K mismatch {qa.shape} vs {qw.shape}")
    return _gemm("b200q_gemm_w8a8", qa, qw, qw.shape[0], K, delta_a, delta_w, zp_w, rowsum_a, bias, out_dtype,
                 epilogue, residual, gate, out)


w4_expand = True            # gemm_w4a8: expand the packed weights once per call when M >= 1024 (False = in-kernel converter)


def pack_w4(codes):
    """int8 codes in [-8,7], [N,K] -> packed uint8 [N, ceil(K/8)*4] (format: csrc/w4.cu)."""
    codes = _rows2d(codes, "pack_w4")
    N, K = codes.shape
    nbytes = ((K + 7) // 8) * 4
    pitch = (nbytes + 15) // 16 * 16                         # TMA global stride: multiple of 16 bytes
    buf = torch.zeros((N, pitch), dtype=torch.uint8, device=codes.device)
    packed = buf[:, :nbytes]
    rc = load().b200q_pack_w4(_ptr(codes), _ld(codes), N, K, _ptr(packed), pitch, _stream())
    _check(rc, "b200q_pack_w4")
    return packed


def gemm_w4a8(qa, qw4, K, delta_a=None, delta_w=None, zp_w=None, rowsum_a=None, bias=None,
              out_dtype=torch.bfloat16, epilogue=EPI_NONE, residual=None, gate=None, out=None):
    """Same contract as gemm_w8a8 with weights packed by pack_w4 (rowsum_a is always required:
    the unsigned-nibble bias is folded through the zero-point term)."""
    if out_dtype == torch.int32:
        raise B200QError("gemm_w4a8: raw accumulators carry the +8 nibble bias; request a dequantised output")
    ws = None
    if w4_expand and qa.shape[0] >= 1024:                     # many-token GEMM: expand the weights once per call (b200q.h)
        ws = _w4_scratch(qa.device, qw4.shape[0] * ((K + 31) // 32) * 32)
    return _gemm("b200q_gemm_w4a8", qa, qw4, qw4.shape[0], K, delta_a, delta_w, zp_w, rowsum_a, bias, out_dtype,
                 epilogue, residual, gate, out, extra=(_ptr(ws),))


# ---------------------------------------------------------------------------------------------
# fused token-local ops
# ---------------------------------------------------------------------------------------------
def ln_mod_quant(x, eps, ln_w=None, ln_b=None, shift=None, scale=None, n_bits=8, quant=True,
                 want_rowsum=True, y_dtype=None):
    """LayerNorm -> affine -> adaLN modulate -> per-token sym quant.  Returns (q, delta, rowsum, y)."""
    x = _rows2d(x, "ln_mod_quant")
    rows, cols = x.shape
    dev = x.device
    q = torch.empty((rows, cols), dtype=torch.int8, device=dev) if quant else None
    delta = torch.empty(rows, dtype=torch.float32, device=dev) if quant else None
    rs = torch.empty(rows, dtype=torch.int32, device=dev) if (quant and want_rowsum) else None
    y = torch.empty((rows, cols), dtype=y_dtype, device=dev) if y_dtype is not None else None

    def vec(t, name):
        if t is None:
            return None
        t = _cuda(t, name).reshape(-1)
        if t.dtype != torch.float32 or t.numel() != cols or not t.is_contiguous():
            t = t.float().contiguous()
        if t.numel() != cols:
            raise B200QError(f"ln_mod_quant: {name} must have {cols} entries")
        return t

    ln_w, ln_b, shift, scale = vec(ln_w, "ln_w"), vec(ln_b, "ln_b"), vec(shift, "shift"), vec(scale, "scale")
    rc = load().b200q_ln_mod_quant(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), _ptr(ln_w), _ptr(ln_b), float(eps),
                                   _ptr(shift), _ptr(scale), int(n_bits), _ptr(q), _ld(q) if q is not None else 0,
                                   _ptr(delta), _ptr(rs), _ptr(y), _DTYPE[y_dtype] if y is not None else F32,
                                   _ld(y) if y is not None else 0, _stream())
    _check(rc, "b200q_ln_mod_quant")
    return q, delta, rs, y


def gate_residual(y, residual, gate=None, out=None):
    """out = residual + y*gate (fp32 residual stream; in place on `residual` by default)."""
    y = _rows2d(y, "gate_residual")
    residual = _rows2d(residual, "gate_residual")
    if residual.dtype != torch.float32:
        raise B200QError("gate_residual: the residual stream is fp32")
    rows, cols = y.shape
    out = residual if out is None else out
    if gate is not None:
        gate = _cuda(gate, "gate").reshape(-1).float().contiguous()
    rc = load().b200q_gate_residual(_ptr(y), _DTYPE[y.dtype], _ld(y), _ptr(gate), _ptr(residual), _ld(residual),
                                    _ptr(out), _ld(out), rows, cols, _stream())
    _check(rc, "b200q_gate_residual")
    return out


def rmsnorm_rope(x, weight, eps, cos=None, sin=None, head_dim=0, head_sq_max=None):
    """RMSNorm over the full dim (+ RoPE when cos/sin [rows, head_dim/2] are given); x bf16/fp16 [rows, cols] (may be a
    strided column slice) -> bf16 [rows, cols].  head_sq_max (fp32 [cols/128], zeroed by the caller): also accumulate the
    per-head maxima of the squared output row norms (the attention's bounded-head classification input)."""
    _cuda(x, "rmsnorm_rope")
    if x.dim() != 2 or x.stride(1) != 1:
        raise B200QError("rmsnorm_rope: expected a row-major 2-D tensor")
    rows, cols = x.shape
    out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device)
    if head_sq_max is not None:
        if head_sq_max.dtype != torch.float32 or head_sq_max.numel() != cols // 128 or not head_sq_max.is_contiguous():
            raise B200QError("rmsnorm_rope: head_sq_max must be a contiguous fp32 tensor of cols / 128 values")
        rc = load().b200q_rmsnorm_rope_stats(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), _ptr(weight), float(eps), _ptr(cos),
                                             _ptr(sin), int(head_dim), _ptr(out), _ld(out), _ptr(head_sq_max), _stream())
        _check(rc, "b200q_rmsnorm_rope_stats")
        return out
    rc = load().b200q_rmsnorm_rope(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), _ptr(weight), float(eps), _ptr(cos),
                                   _ptr(sin), int(head_dim), _ptr(out), _ld(out), _stream())
    _check(rc, "b200q_rmsnorm_rope")
    return out


def rmsnorm_rope_quant(x, weight, eps, cos=None, sin=None, head_dim=128, n_bits=8, want_bf16=False):
    """RMSNorm (+ RoPE) with the attention Q/K quantizer fused in: returns (codes int8 [rows, cols], delta f32
    [rows, cols/128], bf16 [rows, cols] | None).  Per-(token, head) symmetric scales (quant_opensora.py:430-435)."""
    _cuda(x, "rmsnorm_rope_quant")
    if x.dim() != 2 or x.stride(1) != 1:
        raise B200QError("rmsnorm_rope_quant: expected a row-major 2-D tensor")
    rows, cols = x.shape
    q = torch.empty((rows, cols), dtype=torch.int8, device=x.device)
    dq = torch.empty((rows, cols // head_dim), dtype=torch.float32, device=x.device)
    out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    rc = load().b200q_rmsnorm_rope_quant(_ptr(x), _DTYPE[x.dtype], rows, cols, _ld(x), _ptr(weight), float(eps), _ptr(cos),
                                         _ptr(sin), int(head_dim), _ptr(out), _ld(out) if out is not None else 0,
                                         _ptr(q), _ld(q), _ptr(dq), int(n_bits), _stream())
    _check(rc, "b200q_rmsnorm_rope_quant")
    return q, dq, out


# ---------------------------------------------------------------------------------------------
# (c) quantized attention
# ---------------------------------------------------------------------------------------------
def quant_vt(v, n_bits=8):
    """v [Lk, C] -> (vt int8 [C, Lk] view of a 16-byte-pitched buffer, delta f32 [C]): one symmetric scale per
    (head, channel) over all tokens (quant_opensora.py:440-442), codes stored transposed for the P.V product."""
    _cuda(v, "quant_vt")
    if v.dim() != 2 or v.stride(1) != 1:
        raise B200QError("quant_vt: expected a row-major 2-D tensor")
    Lk, C = v.shape
    pitch = (Lk + 15) // 16 * 16
    buf = torch.empty((C, pitch), dtype=torch.int8, device=v.device)
    delta = torch.empty(C, dtype=torch.float32, device=v.device)
    ws = torch.empty(C, dtype=torch.float32, device=v.device)
    rc = load().b200q_quant_vt(_ptr(v), _DTYPE[v.dtype], Lk, C, _ld(v), int(n_bits), _ptr(ws), _ptr(buf), pitch,
                               _ptr(delta), _stream())
    _check(rc, "b200q_quant_vt")
    return buf[:, :Lk], delta


# key splits of attn_bf16 when the caller does not say: None = the library's proposal per shape, 1 = never split (bit-identical
# results for any sharding of the queries; bench.py --verify uses it)
attn_bf16_default_splits = None
attn_bf16_bounded_heads = True      # classify heads by the Cauchy-Schwarz score bound and run bounded heads max-free


def attn_bf16(q, k, v, num_heads, sm_scale=None, out=None, want_lse=False, n_splits=None, qk_sq_max=None):
    """bf16 flash attention (include/b200q.h): q [Lq, H*128], k, v [Lk, H*128] bf16 (row-major, any row pitch that is a
    multiple of 8 elements) -> bf16 [Lq, H*128]; want_lse=True also returns the fp32 [H, Lq] log2-sum-exp.
    n_splits: key splits per work item (None = the library's proposal for this shape; 1 = none).
    qk_sq_max: fp32 [2*H] per-head maxima of |q_i|^2, |k_j|^2 left behind by rmsnorm_rope(head_sq_max=...): skips the
    pre-pass over q and k (b200q_attn_bf16_prenorm)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _cuda(t, n)
        if t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1:
            raise B200QError(f"attn_bf16: {n} must be a row-major 2-D bf16 tensor")
    Lq, D = q.shape
    Lk = k.shape[0]
    H = int(num_heads)
    if k.shape[1] != D or v.shape != k.shape or D % H != 0:
        raise B200QError("attn_bf16: shape mismatch")
    hd = D // H
    sm_scale = hd ** -0.5 if sm_scale is None else float(sm_scale)
    out = torch.empty((Lq, D), dtype=torch.bfloat16, device=q.device) if out is None else out
    lse = torch.empty((H, Lq), dtype=torch.float32, device=q.device) if want_lse else None
    if n_splits is None:
        n_splits = attn_bf16_default_splits
    if n_splits is None:
        n_splits = 1 if want_lse else load().b200q_attn_bf16_splits(Lq, Lk, H)
    norm_ws = torch.empty(2 * H, dtype=torch.float32, device=q.device) if attn_bf16_bounded_heads else None
    part = lse_ws = None
    if n_splits > 1:
        part = torch.empty((n_splits, Lq, D), dtype=torch.bfloat16, device=q.device)
        lse_ws = torch.empty((n_splits, H, Lq), dtype=torch.float32, device=q.device)
    if qk_sq_max is not None and attn_bf16_bounded_heads:
        if qk_sq_max.dtype != torch.float32 or qk_sq_max.numel() != 2 * H or not qk_sq_max.is_contiguous():
            raise B200QError("attn_bf16: qk_sq_max must be a contiguous fp32 tensor of 2 * heads values")
        rc = load().b200q_attn_bf16_prenorm(_ptr(q), _ld(q), _ptr(k), _ld(k), _ptr(v), _ld(v), Lq, Lk, H, hd, sm_scale, _ptr(out),
                                            _ld(out), _ptr(lse), int(n_splits), _ptr(part), _ptr(lse_ws), _ptr(qk_sq_max), _stream())
        _check(rc, "b200q_attn_bf16_prenorm")
        return (out, lse) if want_lse else out
    rc = load().b200q_attn_bf16(_ptr(q), _ld(q), _ptr(k), _ld(k), _ptr(v), _ld(v), Lq, Lk, H, hd, sm_scale, _ptr(out), _ld(out),
                                _ptr(lse), int(n_splits), _ptr(part), _ptr(lse_ws), _ptr(norm_ws), _stream())
    _check(rc, "b200q_attn_bf16")
    return (out, lse) if want_lse else out


def scatter_rows(src_ptrs, dst_ptrs, rows, row_bytes, src_pitch_bytes, dst_pitch_bytes):
    """One launch that copies len(src_ptrs) messages of rows x row_bytes: src_ptrs[i] (row pitch src_pitch_bytes[i]) ->
    dst_ptrs[i] (integer device addresses; the destinations are typically peer-GPU buffers).  The data movement of the
    sequence-parallel attention exchange."""
    n = len(src_ptrs)
    if n != len(dst_ptrs) or n != len(src_pitch_bytes):
        raise B200QError("scatter_rows: src / dst / pitch lists differ in length")
    arr = ctypes.c_void_p * max(n, 1)
    pit = c_int64 * max(n, 1)
    rc = load().b200q_scatter_rows(arr(*src_ptrs), arr(*dst_ptrs), n, int(rows), int(row_bytes), pit(*src_pitch_bytes),
                                   int(dst_pitch_bytes), _stream())
    _check(rc, "b200q_scatter_rows")


def attn_bf16_set_fast(poly_pairs):
    """polynomial pairs of every 8 in the max-free kernel (2..5); -1 = online softmax for every head"""
    if load().b200q_attn_bf16_set_fast(int(poly_pairs)) != 0:
        raise B200QError("b200q_attn_bf16_set_fast: bad value")


def attn_bf16_set_cluster(ctas):
    """1 = single-CTA work items, 2 = CTA pairs with tcgen05.mma.cta_group::2 (default)."""
    if load().b200q_attn_bf16_set_cluster(int(ctas)) != 0:
        raise B200QError("b200q_attn_bf16_set_cluster: 1 or 2")


def attn_bf16_set_variant(variant):
    """Bounded heads: 1 = key-pipelined kernel (default), 0 = two-tile kernel."""
    if load().b200q_attn_bf16_set_variant(int(variant)) != 0:
        raise B200QError("b200q_attn_bf16_set_variant: 0 or 1")


def attn_bf16_set_mode(mode):
    if load().b200q_attn_bf16_set_mode(int(mode)) != 0:
        raise B200QError("b200q_attn_bf16_set_mode: bad mode")


ATTN_MAX_KEYS = 65536        # int32 P.V accumulator bound of one kernel call: Lk * 255 * 127 < 2^31


def _attn_i8_call(qq, dq, kq, dk, vtq, dv, H, hd, sm_scale, out, m, l, pc, ldp, acc):
    Lq, Lk = qq.shape[0], kq.shape[0]
    rc = load().b200q_attn_i8(_ptr(qq), _ld(qq), _ptr(dq), dq.stride(0), dq.stride(1), _ptr(kq), _ld(kq), _ptr(dk),
                              dk.stride(0), dk.stride(1), _ptr(vtq), _ld(vtq), _ptr(dv), Lq, Lk, H, hd, sm_scale,
                              _ptr(out), BF16, _ld(out), _ptr(m), _ptr(l), _ptr(pc), ldp, _ptr(acc),
                              _ld(acc) if acc is not None else 0, _stream())
    _check(rc, "b200q_attn_i8")


def attn_i8(qq, dq, kq, dk, vtq, dv, num_heads, sm_scale=None, out=None, debug=False, max_keys=None):
    """Fused int8 attention (include/b200q.h).  qq [Lq, H*128] int8, dq [Lq, H] f32 (any strides), kq/dk likewise,
    vtq [H*128, Lk] int8 (row pitch multiple of 16), dv [H*128] f32 -> bf16 [Lq, H*128].
    debug=True also returns dict(m, l, p, acc): row max (log2 units) / row sum [H, Lq], P~ codes uint8 [H, Lq, Lk],
    raw int32 P.V accumulators [Lq, H*128].
    Lk > max_keys (default 65,536: the int32 accumulator bound; Wan-14B at 1280x720 has 75,600 keys): the keys are split
    into chunks, each chunk is one kernel call with its own row maximum, and the partial outputs are merged with the
    log-sum-exp weights l_c * 2^(m_c - m) (the flash-attention split-K identity); the attention-map grid is then one step
    per (query row, key chunk)."""
    for t, n in ((qq, "qq"), (dq, "dq"), (kq, "kq"), (dk, "dk"), (vtq, "vtq"), (dv, "dv")):
        _cuda(t, n)
    Lq, D = qq.shape
    Lk = kq.shape[0]
    H = int(num_heads)
    hd = D // H
    if dq.shape != (Lq, H) or dk.shape != (Lk, H) or vtq.shape != (D, Lk) or dv.numel() != D:
        raise B200QError("attn_i8: shape mismatch")
    sm_scale = hd ** -0.5 if sm_scale is None else float(sm_scale)
    dev = qq.device
    out = torch.empty((Lq, D), dtype=torch.bfloat16, device=dev) if out is None else out
    dv = dv.contiguous()
    max_keys = ATTN_MAX_KEYS if max_keys is None else int(max_keys)
    if Lk > max_keys:
        if debug:
            raise B200QError("attn_i8: debug outputs are per kernel call; not available when the keys are split")
        n_chunks = -(-Lk // max_keys)
        step = -(-Lk // n_chunks)
        step = (step + 127) // 128 * 128                       # chunk starts stay 16-byte aligned inside vtq rows
        parts = []
        for k0 in range(0, Lk, step):
            k1 = min(Lk, k0 + step)
            o_c = torch.empty((Lq, D), dtype=torch.bfloat16, device=dev)
            m_c = torch.empty((H, Lq), dtype=torch.float32, device=dev)
            l_c = torch.empty((H, Lq), dtype=torch.float32, device=dev)
            _attn_i8_call(qq, dq, kq[k0:k1], dk[k0:k1], vtq[:, k0:k1], dv, H, hd, sm_scale, o_c, m_c, l_c, None, 0, None)
            parts.append((o_c, m_c, l_c))
        m = torch.stack([p[1] for p in parts]).amax(dim=0)                        # [H, Lq], log2 units
        w = [p[2] * torch.exp2(p[1] - m) for p in parts]                          # l_c * 2^(m_c - m)
        tot = sum(w)
        acc = None
        for (o_c, _, _), w_c in zip(parts, w):
            term = o_c.view(Lq, H, hd).float() * (w_c / tot).t().unsqueeze(-1)
            acc = term if acc is None else acc + term
        out.copy_(acc.view(Lq, D))
        return out
    m = l = pc = acc = None
    ldp = 0
    if debug:
        m = torch.empty((H, Lq), dtype=torch.float32, device=dev)
        l = torch.empty((H, Lq), dtype=torch.float32, device=dev)
        ldp = (Lk + 127) // 128 * 128
        pc = torch.zeros((H, Lq, ldp), dtype=torch.uint8, device=dev)
        acc = torch.empty((Lq, D), dtype=torch.int32, device=dev)
    _attn_i8_call(qq, dq, kq, dk, vtq, dv, H, hd, sm_scale, out, m, l, pc, ldp, acc)
    if debug:
        return out, dict(m=m, l=l, p=pc[:, :, :Lk], acc=acc)
    return out


def gemm_set_cluster(mode):
    """0 auto, 1 single-CTA tiles, 2 force 2-CTA multicast clusters (results identical; scheduling knob)."""
    rc = load().b200q_gemm_set_cluster(int(mode))
    if rc != 0:
        raise B200QError("b200q_gemm_set_cluster: bad mode")
