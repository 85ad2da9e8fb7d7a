"""QuantizedLinear — same constructor, attributes and forward contract as the reference layer
(ViDiT-Q/quant_utils/qdiff/base/quant_layer.py:8-74), executed as REAL integer arithmetic:

    reference:  y = F.linear( deq(quant(x)), deq(quant(W)), bias )          (fake-quant, fp)
    here     :  qa,da,rowsum = b200q.quant_rows(x)            (one kernel, per-token, no host sync)
                y = b200q.gemm_w8a8(qa, qW, da, dW, zpW, rowsum, bias)      (tcgen05 int8 GEMM)
              = da[m]*dW[n]*( sum_k qa*qW + zpW[n]*sum_k qa ) + bias[n]      — algebraically identical

`weight` keeps the dequantised fake-quant weight, `fp_weight`/`fp_module` the original (the reference's
checkpoints, quant_param_dict and `quantize_and_save_weight_` read these); the int8/int4 codes the kernel
consumes live in `qweight` (+ `qweight_packed` for 4-bit) and are rebuilt whenever the weight quantizer
changes (bitwidth_refactor, load_quant_param_dict)."""
import torch
import torch.nn.functional as F

import b200q
from qdiff.base.base_quantizer import DynamicQuantizer, StaticQuantizer
from qdiff.base.mixed_precision_quantizer import MixedPrecisionDynamicQuantizer, MixedPrecisionStaticQuantizer

try:
    from omegaconf import ListConfig
except ImportError:  # pragma: no cover
    ListConfig = ()


def _is_list(v):
    return isinstance(v, (list, tuple)) or bool(ListConfig and isinstance(v, ListConfig))


class ActPlan:
    """What a layer does to its activations in front of the per-token quantizer, in the factored form the fused kernel
    b200q_had_quant_rows takes (include/b200q.h):  y = (x * colscale) . (H_K (x) H_{2^log2w}),
    colscale = channel_mask * sign / sqrt(n).  `mask` / `sign` keep the un-folded factors for the int-weight checkpoint."""

    def __init__(self, n, colscale, hadK, K, log2w, mask=None, sign=None):
        self.n, self.colscale, self.hadK, self.K, self.log2w, self.mask, self.sign = n, colscale, hadK, K, log2w, mask, sign

    @staticmethod
    def rotation(n, sign, mask, device):
        """R = diag(sign) . H_n / sqrt(n), optionally after the smooth scale `mask`; None if the kernel cannot serve n."""
        from qdiff.quarot.quarot_utils import hadamard_kernel_plan
        kp = hadamard_kernel_plan(n)
        if kp is None:
            return None
        K, w, H = kp
        cs = sign.detach().double().cpu().reshape(-1) / (n ** 0.5)
        if mask is not None:
            cs = cs * mask.detach().double().cpu().reshape(-1)
        hadK = None if H is None else H.to(torch.float32).reshape(-1).contiguous().to(device)
        return ActPlan(n, cs.to(torch.float32).contiguous().to(device), hadK, K, w,
                       None if mask is None else mask.detach().float().reshape(-1).cpu(), sign.detach().float().reshape(-1).cpu())

    @staticmethod
    def scale_only(mask, device):
        n = mask.numel()
        if n % 128 != 0:
            return None
        m = mask.detach().float().reshape(-1)
        return ActPlan(n, m.contiguous().to(device), None, 1, 0, m.cpu(), None)

    def to(self, device):
        return ActPlan(self.n, self.colscale.to(device), None if self.hadK is None else self.hadK.to(device), self.K,
                       self.log2w, self.mask, self.sign)

    def quantize(self, x2d, n_bits=8, want_rowsum=True):
        """-> (codes int8 [rows, n], delta f32 [rows], rowsum int32 [rows] | None)"""
        q, d, rs, _ = b200q.had_quant_rows(x2d, self.colscale, self.hadK, self.K, self.log2w, n_bits, want_rowsum=want_rowsum)
        return q, d, rs


class QuantizedLinear(torch.nn.Linear):
    """Static per-out-channel weight quantization + dynamic per-token activation quantization."""

    def __init__(self, in_features: int, out_features: int, bias: bool, device: None, quant_config: dict,
                 fp_module: torch.nn.Linear) -> None:
        super().__init__(in_features, out_features, bias, device)
        self.fp_module = fp_module
        self.q_cfg = quant_config
        self.w_quantizer = None
        self.a_quantizer = None
        self.module_name = None
        self._int_state = None            # (codes int8 [N,K], packed uint8 | None, delta f32 [N], zp f32 [N] | None, n_bits)

        w_cfg = quant_config.get("weight", None)
        if w_cfg is not None:
            mixed = _is_list(w_cfg["n_bits"])
            self.w_quantizer = (MixedPrecisionStaticQuantizer if mixed else StaticQuantizer)(w_cfg)
            # weights are quantized once, from the FP module; bias stays FP (quant_layer.py:38-41)
            self.weight.data = self.w_quantizer(fp_module.weight)
            self.w_quantizer.init_done = True
        else:
            self.weight.data = fp_module.weight

        self.fp_weight = self.fp_module.weight
        self.bias = fp_module.bias

        a_cfg = quant_config.get("act", None)
        if a_cfg is not None:
            mixed = _is_list(a_cfg["n_bits"])
            self.a_quantizer = (MixedPrecisionDynamicQuantizer if mixed else DynamicQuantizer)(a_cfg)

        self.use_kernel = True     # the reference keeps this False and simulates; here the kernel IS the path
        self.quant_mode = True     # False -> run the original FP module (quant_layer.py:61-62)

    # ---- integer weight state ----------------------------------------------------------------------
    def invalidate_int_weight(self):
        self._int_state = None

    def int_weight_state(self, device):
        """int8 codes (+ 4-bit packing), fp32 delta / zero_point vectors on `device`, built lazily from the
        FP weight with the weight quantizer's current parameters (the codes the reference exports in
        quantize_and_save_weight_, examples/Wan2.1/wan/quant_wanx_cuda.py:39-55, without its fp16 detour)."""
        wq = self.w_quantizer
        st = self._int_state
        if st is not None and st["device"] == device and st["n_bits"] == wq.n_bits and st["delta_id"] is wq.delta:
            return st
        w = self._weight_for_codes().detach().to(device)
        delta, zp = wq.params_f32(device) if hasattr(wq, "params_f32") else (
            wq.delta.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous(),
            wq.zero_point.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous())
        codes, _ = b200q.quant_rows_static(w, delta, zp, wq.n_bits, wq.sym)
        K = codes.shape[1]
        if K % 16 != 0:                    # TMA row pitch: pad the K axis with zero codes once, view back to K
            padded = torch.zeros((codes.shape[0], (K + 15) // 16 * 16), dtype=torch.int8, device=device)
            padded[:, :K] = codes
            codes = padded[:, :K]
        packed = b200q.pack_w4(codes) if wq.n_bits <= 4 else None
        st = dict(device=device, n_bits=wq.n_bits, delta_id=wq.delta, codes=codes, packed=packed, delta=delta,
                  zp=None if wq.sym else zp)
        self._int_state = st
        return st

    def _weight_for_codes(self):
        # plain layer: codes come from the FP weight; SmoothQuant/QuaRot/ViDiT-Q subclasses override this
        return self.fp_module.weight

    # ---- forward ---------------------------------------------------------------------------------------
    def _out_dtype(self, x):
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype("cuda")      # bf16 under the pipeline's autocast (text2video.py:213)
        return x.dtype if x.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32

    def _prepare_activation(self, x2d):
        return x2d        # hook for the smooth/rotate variants (unfused form: torch ops in front of the quantizer)

    def _act_plan(self, device):
        """ActPlan for the fused smooth/rotate + quantize kernel, or None: plain layer / transform the kernel cannot
        serve (then `_prepare_activation` runs in front of the row quantizer)."""
        return None

    def forward(self, x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        """x: [B, N_token, C] (any leading shape is accepted) -> [B, N_token, C_out]"""
        if not self.quant_mode:
            return self.fp_module(x, *args, **kwargs)
        lead = x.shape[:-1]
        x2d = x.reshape(-1, x.shape[-1])
        if self.w_quantizer is None or self.a_quantizer is None:
            # weight-only / activation-only configs are not an integer GEMM: run the dequantised operands
            # through the library GEMM exactly as the reference does
            if self.a_quantizer is not None:
                x2d = self.a_quantizer(self._prepare_activation(x2d))
            return F.linear(x2d, self.weight.to(x2d.dtype), self.bias).reshape(*lead, self.out_features)
        if not self.a_quantizer.sym:
            raise NotImplementedError(
                "asymmetric activation quantization needs the zp_a*colsum(W) epilogue term, which libb200q does not "
                "implement; the Wan2.1 configs use symmetric per-token activations (quant_configs/config.yaml:17-18)")
        st = self.int_weight_state(x2d.device)
        plan = self._act_plan(x2d.device) if x2d.shape[1] % 16 == 0 else None
        if plan is not None:
            qa, da, rowsum = plan.quantize(x2d, self.a_quantizer.n_bits)
        else:
            x2d = self._prepare_activation(x2d)
            qa, da, _, rowsum = self.a_quantizer.quantize_int8(x2d, want_rowsum=True)
        K = x2d.shape[1]
        if K % 16 != 0:
            padded = torch.zeros((qa.shape[0], (K + 15) // 16 * 16), dtype=torch.int8, device=qa.device)
            padded[:, :K] = qa
            qa = padded[:, :K]
        out_dtype = self._out_dtype(x)
        if st["packed"] is not None:
            y = b200q.gemm_w4a8(qa, st["packed"], K, da, st["delta"], st["zp"], rowsum, self.bias, out_dtype=out_dtype)
        else:
            y = b200q.gemm_w8a8(qa, st["codes"], da, st["delta"], st["zp"], rowsum, self.bias, out_dtype=out_dtype)
        return y.reshape(*lead, self.out_features)
