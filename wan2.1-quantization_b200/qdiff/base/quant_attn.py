"""Post-softmax attention-map quantizers (mirror of ViDiT-Q/quant_utils/qdiff/base/quant_attn.py).

Only the per-key-column group ('row' in the OpenSORA class :168-174, 'column' in the CogVideoX class
:47-50) is in scope: one dynamic scale per KEY column taken over all queries of a head.  The 'block'
group needs reorder tables and token grids hard-coded for OpenSORA/CogVideoX (:52-59, :176-183) and
is out of scope for Wan2.1.

This module quantizes a MATERIALISED attention map [B,H,Lq,Lk] and exists for parity at small L and for
API compatibility; the production path is the fused int8 attention kernel, which never materialises P."""
import torch

import b200q
from qdiff.base.base_quantizer import DynamicQuantizer


class _AttnMapQuantizer(torch.nn.Module):
    _column_group_name = "row"

    def __init__(self, quant_config, cross_attn=False):
        super().__init__()
        self.quant_config = quant_config
        self.group = quant_config.attn.attn_map.group
        section = quant_config["cross_attn"] if cross_attn else quant_config["attn"]
        self.attn_map_quantizer = DynamicQuantizer(section["attn_map"])
        self.cross_attn = cross_attn
        self.mixed_precision_cfg = None
        self.i_block = None
        self.split_range = None
        self.quant_mode = True

    def forward(self, x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        """x: softmax output [B, H, Lq, Lk] -> fake-quantised map of the same shape."""
        if not self.quant_mode:
            return x
        if self.group != self._column_group_name:
            raise NotImplementedError(
                f"attn_map.group={self.group!r}: only the per-key-column group ({self._column_group_name!r}) is "
                "supported on the Wan2.1 path")
        B, H, Lq, Lk = x.shape
        cols = x.permute(0, 1, 3, 2).reshape(-1, Lq)          # one row per key column (quant_attn.py:169)
        deq = self.attn_map_quantizer(cols)
        return deq.reshape(B, H, Lk, Lq).permute(0, 1, 3, 2)


class QuantizedAttentionMapOpenSORA(_AttnMapQuantizer):
    """quant_attn.py:118-174 ('row' group)."""
    _column_group_name = "row"


class QuantizedAttentionMap(_AttnMapQuantizer):
    """quant_attn.py:8-50 ('column' group) — same arithmetic, CogVideoX naming."""
    _column_group_name = "column"

    def __init__(self, quant_config):
        super().__init__(quant_config, cross_attn=False)
