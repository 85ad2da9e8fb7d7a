"""Module-tree surgery and quant-parameter (de)serialisation with the reference's function names and
keyword contracts (ViDiT-Q/quant_utils/qdiff/base/quant_model.py), so QuantWanModel
(examples/Wan2.1/wan/quant_wanx.py:85-133) can call them unchanged through apply_func_to_submodules."""
import logging
import re

import torch
import torch.nn as nn

from qdiff.base.base_quantizer import BaseQuantizer, DynamicQuantizer, StaticQuantizer  # noqa: F401
from qdiff.base.quant_layer import QuantizedLinear
from qdiff.utils import apply_func_to_submodules

logger = logging.getLogger(__name__)

# method section in the YAML -> (module path, class name); the LAST matching section wins, as in the
# reference's three consecutive `if` blocks (quant_model.py:21-53)
_METHOD_LAYERS = (
    ("smooth_quant", "qdiff.smooth_quant.sq_quant_layer", "SQQuantizedLinear"),
    ("quarot", "qdiff.quarot.quarot_quant_layer", "QuarotQuantizedLinear"),
    ("viditq", "qdiff.viditq.viditq_quant_layer", "ViDiTQuantizedLinear"),
)


def _select_layer_type(quant_config, full_name):
    chosen = QuantizedLinear
    for section, module_path, cls_name in _METHOD_LAYERS:
        cfg = quant_config.get(section, None)
        if cfg is None:
            continue
        if re.search(re.compile(cfg.layer_name_regex), full_name):
            import importlib
            chosen = getattr(importlib.import_module(module_path), cls_name)
            logger.info("setting %s for layer %s", section, full_name)
    return chosen


def quant_layer_refactor_(submodule, name, parent_module, quant_config, full_name, remain_fp_regex):
    """Replace one nn.Linear by its quantized counterpart unless `remain_fp_regex` matches its dotted name
    (quant_model.py:15-74)."""
    layer_type = _select_layer_type(quant_config, full_name)
    if remain_fp_regex is not None and re.compile(remain_fp_regex).search(full_name):
        logger.info("remain %s quant as FP due to fp_regex", full_name)
        return
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    new_layer = layer_type(submodule.in_features, submodule.out_features, submodule.bias is not None, device,
                           quant_config, submodule)
    setattr(parent_module, name, new_layer)
    new_layer.module_name = full_name
    for qz in (new_layer.w_quantizer, new_layer.a_quantizer):
        if qz is not None:
            qz.module_name = full_name


def bitwidth_refactor_(submodule, name, parent_module, quant_config, full_name):
    """mixed_precision.{weight,act}.layer_name_regex = [fp16_regex, regex_for_bits[0], regex_for_bits[1], ...]
    (quant_model.py:76-105): index 0 switches the layer back to FP, index i>0 selects n_bits[i-1]."""
    mp = quant_config.mixed_precision
    for kind, regex_list in (("W", mp.weight.layer_name_regex), ("A", mp.act.layer_name_regex)):
        quantizer = submodule.w_quantizer if kind == "W" else submodule.a_quantizer
        for idx, layer_regex in enumerate(regex_list):
            if len(layer_regex) == 0:
                continue
            if not re.search(re.compile(layer_regex), full_name):
                continue
            if idx == 0:
                submodule.quant_mode = False
                logger.info("[Mixed Precision] set the %s %s as FP16", full_name, kind)
            else:
                quantizer.bitwidth_refactor(idx - 1)
                logger.info("[Mixed Precision] set the %s %s as %s bit", full_name, kind, quantizer.bitwidth_list[idx - 1])
    if hasattr(submodule, "invalidate_int_weight"):
        submodule.invalidate_int_weight()
        if submodule.w_quantizer is not None and submodule.quant_mode:
            # keep `weight` (the fake-quant view of W) consistent with the selected bit-width
            submodule.weight.data = submodule.w_quantizer(submodule._weight_for_codes())


def load_quant_param_dict_(submodule, full_name, parent_module, quant_param_dict, model):
    """quant_model.py:138-160 — `submodule` is a quantizer, `full_name` e.g. 'blocks.0.self_attn.q.w_quantizer'."""
    entry = quant_param_dict[full_name]
    submodule.delta = entry["delta"]
    submodule.zero_point = entry["zero_point"]
    has_mask, has_rot = hasattr(parent_module, "channel_mask"), hasattr(parent_module, "rotation_matrix")
    if has_mask and has_rot:            # ViDiT-Q
        parent_module.get_rotation_matrix()
        parent_module.channel_mask = entry["channel_mask"]
        parent_module.update_quantized_weight_rotated_and_scaled()
    elif has_rot:                       # QuaRot
        parent_module.get_rotation_matrix()
        parent_module.update_quantized_weight_rotated()
    elif has_mask:                      # SmoothQuant
        parent_module.channel_mask = entry["channel_mask"]
        parent_module.update_quantized_weight_scaled()
    if hasattr(parent_module, "invalidate_int_weight"):
        parent_module.invalidate_int_weight()
    model.quant_param_dict[full_name] = entry


def save_quant_param_dict_(submodule, full_name, parent_module, model):
    """quant_model.py:161-172 — schema {'<layer>.w_quantizer': {'delta':[C_out,1], 'zero_point':[C_out,1],
    ('channel_mask':[C_in]), ('rotation_matrix': None)}, '<layer>.a_quantizer': {...}}"""
    entry = {"delta": submodule.delta, "zero_point": submodule.zero_point}
    if hasattr(parent_module, "channel_mask"):
        entry["channel_mask"] = parent_module.channel_mask
    if hasattr(parent_module, "rotation_matrix"):
        entry["rotation_matrix"] = None      # large and identical across layers: regenerated on load
    model.quant_param_dict[full_name] = entry


def set_init_done_(submodule):
    submodule.init_done = True


class QuantModel(nn.Module):
    """Template (quant_model.py:182-234): subclass your model, set `quant_config`, call these."""

    def __init__(self, quant_config: dict, **kwargs) -> None:
        super().__init__()
        self.q_cfg = quant_config
        self.quant_config = quant_config
        self.quant_param_dict = {}

    def quant_layer_refactor(self):
        apply_func_to_submodules(self, class_type=nn.Linear, function=quant_layer_refactor_, name=None,
                                 parent_module=None, quant_config=self.quant_config, full_name=None,
                                 remain_fp_regex=self.quant_config.get("remain_fp_regex", None))

    def save_quant_param_dict(self):
        apply_func_to_submodules(self, class_type=BaseQuantizer, function=save_quant_param_dict_, full_name=None,
                                 parent_module=None, model=self)

    def load_quant_param_dict(self, quant_param_dict):
        apply_func_to_submodules(self, class_type=BaseQuantizer, function=load_quant_param_dict_, full_name=None,
                                 parent_module=None, quant_param_dict=quant_param_dict, model=self)

    # the template's own spellings (quant_model.py:204-214)
    save_quant_params_dict = save_quant_param_dict
    load_quant_params_dict = load_quant_param_dict

    def set_init_done(self):
        apply_func_to_submodules(self, class_type=BaseQuantizer, function=set_init_done_)

    def bitwidth_refactor(self):
        apply_func_to_submodules(self, class_type=QuantizedLinear, function=bitwidth_refactor_, name=None,
                                 parent_module=None, quant_config=self.quant_config, full_name=None)

    def forward(self, x, *args, **kwargs):
        raise NotImplementedError("should be implemented in subclass.")
