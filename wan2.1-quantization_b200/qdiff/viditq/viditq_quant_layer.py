"""ViDiT-Q layer = per-channel smooth scale, then randomized Hadamard rotation
(ViDiT-Q/quant_utils/qdiff/viditq/viditq_quant_layer.py:8-73; selected for every layer by the shipped YAML,
quant_configs/config.yaml:19-21), integer execution:

    weights : W / mask -> fake-quant -> rotate -> quantised again (the reference quantises twice, :42-50; kept so that
              `weight`, `delta`, `zero_point` match the reference's)                                   (offline)
    forward : (x * mask) @ R -> per-token quantizer -> tcgen05 int8 GEMM                               (:52-73)"""
import torch

from qdiff.base.quant_layer import QuantizedLinear
from qdiff.quarot.quarot_quant_layer import RotationMixin


class ViDiTQuantizedLinear(RotationMixin, QuantizedLinear):
    def __init__(self, in_features, out_features, bias, device, quant_config, fp_module):
        super().__init__(in_features, out_features, bias, device, quant_config, fp_module)
        self.alpha = quant_config.viditq.alpha
        self.channel_mask = None          # assigned outside, during PTQ (ptq_wanx.py:334-344)
        self.rotation_matrix = None
        self._rot_plan = None
        self._rotated_weight = None

    def get_channel_mask(self, act_mask):
        weight_mask = self.fp_module.weight.abs().max(dim=0)[0]                     # [C_in]
        channel_mask = (weight_mask.abs() ** self.alpha) / (act_mask.abs() ** (1 - self.alpha))
        self.channel_mask = channel_mask
        assert not torch.isnan(channel_mask).any(), "nan exists in channel mask"

    def update_quantized_weight_rotated_and_scaled(self):
        assert self.channel_mask is not None
        C_out, C_in = self.fp_module.weight.shape
        w = self.fp_module.weight
        self.w_quantizer.init_done = False
        self.weight.data = self.w_quantizer(w / self.channel_mask.reshape([1, C_in]).to(w.device))
        self._rotated_weight = torch.matmul(self.weight.data.double(), self.rotation_matrix.to(w.device)).float()
        self.weight.data = self.w_quantizer(self._rotated_weight)
        self.w_quantizer.init_done = True
        self.invalidate_int_weight()

    def _weight_for_codes(self):
        return self._rotated_weight if self._rotated_weight is not None else self.fp_module.weight

    def _prepare_activation(self, x2d):
        if self.channel_mask is None:
            raise RuntimeError("ViDiTQuantizedLinear: channel_mask not set (run PTQ or load_quant_param_dict first)")
        return self._rotate(x2d * self.channel_mask.to(device=x2d.device, dtype=x2d.dtype).reshape(1, -1))

    def _act_plan(self, device):
        if self.channel_mask is None:
            raise RuntimeError("ViDiTQuantizedLinear: channel_mask not set (run PTQ or load_quant_param_dict first)")
        return self._rotation_act_plan(device, self.channel_mask)
