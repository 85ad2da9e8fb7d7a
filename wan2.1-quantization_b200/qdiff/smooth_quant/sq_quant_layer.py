"""SmoothQuant layer (ViDiT-Q/quant_utils/qdiff/smooth_quant/sq_quant_layer.py:5-68), integer execution:

    channel_mask = |W|max(col)^alpha / |X|max(col)^(1-alpha)                          (:27-34, set during PTQ)
    weights      : W / mask re-quantised per out-channel (offline)                      (:36-44)
    forward      : (x * mask) -> per-token quantizer -> tcgen05 int8 GEMM             (:46-68)

The per-channel multiply is one elementwise pass in front of the row quantizer (fusing it into quant_rows_kernel is
the remaining step of SURVEY §8 f-2)."""
import torch

from qdiff.base.quant_layer import ActPlan, QuantizedLinear


class SQQuantizedLinear(QuantizedLinear):
    def __init__(self, in_features, out_features, bias, device, quant_config, fp_module):
        super().__init__(in_features, out_features, bias, device, quant_config, fp_module)
        self.alpha = quant_config.smooth_quant.alpha
        self.channel_mask = None          # assigned outside, during PTQ (ptq_wanx.py:336-344)
        self._scaled_weight = None

    def get_channel_mask(self, act_mask):
        weight_mask = self.fp_module.weight.abs().max(dim=0)[0]                     # [C_in]
        channel_mask = (weight_mask.abs() ** self.alpha) / (act_mask.abs() ** (1 - self.alpha))
        self.channel_mask = channel_mask
        assert not torch.isinf(self.channel_mask).any().item(), "inf exists in channel_mask"

    def update_quantized_weight_scaled(self):
        assert self.channel_mask is not None
        C_out, C_in = self.fp_module.weight.shape
        self.w_quantizer.init_done = False
        self._scaled_weight = self.fp_module.weight / self.channel_mask.reshape([1, C_in]).to(self.fp_module.weight.device)
        self.weight.data = self.w_quantizer(self._scaled_weight)
        assert not torch.isnan(self.weight.data).any().item(), "nan exists in weight"
        self.w_quantizer.init_done = True
        self.invalidate_int_weight()

    def _weight_for_codes(self):
        return self._scaled_weight if self._scaled_weight is not None else self.fp_module.weight

    def _prepare_activation(self, x2d):
        if self.channel_mask is None:
            raise RuntimeError("SQQuantizedLinear: channel_mask not set (run PTQ or load_quant_param_dict first)")
        return x2d * self.channel_mask.to(device=x2d.device, dtype=x2d.dtype).reshape(1, -1)

    def _act_plan(self, device):
        if self.channel_mask is None:
            raise RuntimeError("SQQuantizedLinear: channel_mask not set (run PTQ or load_quant_param_dict first)")
        cached = getattr(self, "_act_plan_cache", None)
        if cached is None or cached[0] is not self.channel_mask or cached[1] != device:
            cached = self._act_plan_cache = (self.channel_mask, device, ActPlan.scale_only(self.channel_mask, device))
        return cached[2]
