"""Quantizers with the reference's interface (ViDiT-Q/quant_utils/qdiff/base/base_quantizer.py) whose
arithmetic runs in libb200q:

  BaseQuantizer      :13-41   n_bits / sym / n_levels, `delta` / `zero_point` buffers, init_done
  StaticQuantizer    :43-99   offline per-row parameters (weights: one row per out-channel)
  DynamicQuantizer   :101-206 online per-row parameters (activations: one row per token)

Beyond the reference surface each quantizer has `quantize_int8(x)`, the real-integer entry the
quantized linear uses: int8 codes + fp32 scales straight from the kernel, no dequantise round trip.
Deviations, all on paths where the reference stops in `ipdb`: a zero-range row gets delta=1e-8
(asym) instead of a debugger prompt; NaN inputs are not asserted on (that assert is a host sync).
"""
import logging

import torch
import torch.nn as nn

import b200q

try:
    from omegaconf import ListConfig
except ImportError:  # pragma: no cover - the compat shim is put on sys.path by the caller
    ListConfig = ()

logger = logging.getLogger(__name__)


def _levels(n_bits, sym):
    # base_quantizer.py:32
    return 2 ** (n_bits - 1) - 1 if sym else 2 ** n_bits


def _on_cuda(x):
    """Kernels need device memory; parameters that still live on the host (PTQ before FSDP wrap,
    ptq_wanx.py:315-320) are staged through the current CUDA device.  No GPU -> loud failure."""
    if x.is_cuda:
        return x, None
    if not torch.cuda.is_available():
        raise b200q.B200QError("qdiff quantizers run on libb200q (CUDA); no GPU is visible and there is no CPU path")
    return x.to("cuda"), x.device


class BaseQuantizer(nn.Module):
    def __init__(self, quant_config):
        super().__init__()
        self.n_bits = quant_config["n_bits"]
        self.sym = quant_config.get("sym", False)
        if isinstance(self.n_bits, list):
            raise AssertionError("when multiple n_bits are adopted, use the MixedPrecisionBaseQuantizer")
        self.register_buffer("delta", None)
        self.register_buffer("zero_point", None)
        if not (ListConfig and isinstance(self.n_bits, ListConfig)):
            self.n_levels = _levels(self.n_bits, self.sym)
        self.init_done = False
        self.module_name = None

    # ---- shared kernel plumbing ------------------------------------------------------------
    def _compute(self, x, dynamic):
        """Run the row quantizer kernel; record delta/zero_point ([G,1], x's device) and row statistics."""
        assert x.dim() == 2, "quantizers take [N_group, -1]"
        xc, home = _on_cuda(x.detach())
        q, delta, zp, rowsum, smax, smin = b200q.quant_rows(xc, self.n_bits, self.sym, dynamic, want_stats=True)
        back = (lambda t: t.to(home)) if home is not None else (lambda t: t)
        self.delta = back(delta).unsqueeze(-1).to(x.dtype)
        self.zero_point = back(zp).unsqueeze(-1).to(x.dtype)
        # the kernel derives the codes from fp32 parameters; for 16-bit inputs `delta` / `zero_point` above are rounded
        # to x.dtype (the reference's buffers have x's dtype), so the exact pair is kept for the integer path
        # (QuantizedLinear.int_weight_state): codes == the codes behind the fake-quant weight, bit for bit
        self._delta_f32, self._zp_f32, self._f32_of = back(delta), back(zp), self.delta
        if self.sym:
            self.x_absmax = back(smax)
        else:
            self.x_max, self.x_min = back(smax), back(smin)
        return q, delta, zp, rowsum, home

    def params_f32(self, device):
        """(delta, zero_point) fp32 [G] on `device`: the exact kernel parameters while `delta` is still the tensor they
        produced, else the (possibly loaded) buffers."""
        if getattr(self, "_f32_of", None) is self.delta and self.delta is not None:
            d, z = self._delta_f32, self._zp_f32
        else:
            d, z = self.delta, self.zero_point
        return (d.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous(),
                z.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous())

    def _apply_static(self, x):
        xc, home = _on_cuda(x.detach())
        d, z = self.params_f32(xc.device)
        q, _ = b200q.quant_rows_static(xc, d, z, self.n_bits, self.sym)
        return q, home

    @staticmethod
    def _codes_as(q, like, home):
        # integer codes as a float tensor, the reference's return convention (base_quantizer.py:66-68)
        ones = torch.ones(q.shape[0], dtype=torch.float32, device=q.device)
        out = b200q.dequant_rows(q, ones, None, like.dtype if like.dtype.is_floating_point else torch.float32)
        return out.to(home) if home is not None else out

    def _dequant(self, q, like, home):
        dev = q.device
        # fake-quant output = (q + zp) * delta with the buffers as the reference holds them (x's dtype)
        out = b200q.dequant_rows(q, self.delta.to(dev), self.zero_point.to(dev), like.dtype)
        return out.to(home) if home is not None else out

    def forward(self, x: torch.Tensor):
        raise NotImplementedError("should be implemented in subclass.")

    def init_quant_params(self, x):
        raise NotImplementedError("should be implemented in subclass.")


class StaticQuantizer(BaseQuantizer):
    """Input [Group, -1]; parameters fixed offline by `init_quant_params` (base_quantizer.py:43-99)."""

    def __init__(self, quant_config):
        super().__init__(quant_config)
        if self.sym:
            self.x_absmax = None
        else:
            self.x_max = None
            self.x_min = None

    def init_quant_params(self, x):
        """base_quantizer.py:70-99: delta / zero_point come from THIS call's rows; the stored statistics x_absmax /
        x_max / x_min are running maxima / minima over calls (:76, :84, :88)."""
        prev = (self.x_absmax,) if self.sym else (self.x_max, self.x_min)
        self._compute(x, dynamic=False)
        if prev[0] is not None and prev[0].shape == (self.x_absmax if self.sym else self.x_max).shape:
            if self.sym:
                self.x_absmax = torch.max(self.x_absmax, prev[0].to(self.x_absmax.device))
            else:
                self.x_max = torch.max(self.x_max, prev[0].to(self.x_max.device))
                self.x_min = torch.min(self.x_min, prev[1].to(self.x_min.device))
        if not bool(torch.all(self.delta > 1.0e-6)):
            logger.warning("unexpected small delta exists in %s (min %.3e)", self.module_name, float(self.delta.min()))

    def quantize_int8(self, x):
        """-> int8 codes [G, C] on the CUDA device (base_quantizer.py:63-68 without the float detour)."""
        if self.init_done is not True:
            self.init_quant_params(x)
        return self._apply_static(x)[0]

    def quantize(self, x: torch.Tensor):
        if self.init_done is not True:
            self.init_quant_params(x)
        q, home = self._apply_static(x)
        return self._codes_as(q, x, home)

    def forward(self, x: torch.Tensor):
        if self.init_done is not True:
            self.init_quant_params(x)
        q, home = self._apply_static(x)
        return self._dequant(q, x, home)


class DynamicQuantizer(BaseQuantizer):
    """Input [Group, -1]; parameters recomputed per call (base_quantizer.py:101-162)."""

    def __init__(self, quant_config):
        super().__init__(quant_config)

    def quantize_int8(self, x, want_rowsum=True):
        """-> (codes int8 [G,C], delta f32 [G], zero_point f32 [G], rowsum int32 [G]) — the hot-path entry:
        one kernel, no host sync (the reference syncs twice per call, :113,:124)."""
        assert x.dim() == 2
        return b200q.quant_rows(x, self.n_bits, self.sym, True, want_rowsum=want_rowsum)

    def quantize(self, x: torch.Tensor):
        q, _, _, _, home = self._compute(x, dynamic=True)
        return self._codes_as(q, x, home)

    def forward(self, x: torch.Tensor):
        q, _, _, _, home = self._compute(x, dynamic=True)
        return self._dequant(q, x, home)

    def forward_with_quant_params(self, x, delta, mixed_precision=None):
        """Block-wise attention-map variant with a precomputed `delta` of x's shape
        (base_quantizer.py:164-206).  Only reachable from the OpenSORA/CogVideoX 'block' attn-map
        group, whose token grids are hard-coded (quant_attn.py:56-59) — out of scope for Wan
        (SURVEY §2.1 #5)."""
        raise NotImplementedError(
            "forward_with_quant_params serves the OpenSORA 'block' attention-map group, which is out of scope "
            "for the Wan2.1 path; use the 'row' group (QuantizedAttentionMapOpenSORA)")
