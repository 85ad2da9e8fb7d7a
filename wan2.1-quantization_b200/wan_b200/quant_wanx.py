"""QuantWanModel / QuantWanMixin — the reference's `QuantWanModel`
(ViDiT-Q/examples/Wan2.1/wan/quant_wanx.py:28-228).

`QuantWanModel` below is concrete: the mixin over `wan_b200.wan_model.WanModelFP` (a self-contained module tree with
WanModel's parameter names, constructor arguments and call convention), with the reference's constructor signature
(`quant_config=` keyword, :36-54) and `from_pretrained(ckpt_dir, quant_config=...)` (quant_generate.py:358-360), so the
reference's sequence

    model = QuantWanModel.from_pretrained(ckpt_dir, quant_config=cfg); model.quant_layer_refactor()
    model.load_quant_param_dict(params); model.quantize_and_save_weight(path)
    model.hardware_forward_refactor(load_path=path, seq_len=L); model.set_init_done(); model(x, t, context, seq_len)

(quant_generate.py:355-420) runs against this package alone.  Where the reference's own `wan` package is importable the
mixin composes with its model instead:

    from wan.modules.model import WanModel                      # the reference's model, unchanged
    from wan_b200.quant_wanx import QuantWanMixin
    class QuantWanModel(QuantWanMixin, WanModel): ...

  convert_quant / quant_layer_refactor / save_quant_param_dict / load_quant_param_dict / set_init_done /
  bitwidth_refactor                      same bodies as quant_wanx.py:80-135 (apply_func_to_submodules over the mirror)
  quantize_and_save_weight(path)         quant_wanx.py:137-185  -> wan_b200.int_checkpoint (fp32 scales, codes == the
                                         fake-quant codes; the reference re-derives them from fp16-rounded operands)
  hardware_forward_refactor(path, seq_len)
                                         quant_wanx.py:188-228 swaps every block for a `viditq_extension` block; here the
                                         whole DiT forward is replaced by the integer runtime (`WanDiTQ`, replayed as a
                                         CUDA graph), with the reference's call convention forward(x, t, context, seq_len).
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn

from qdiff.base.base_quantizer import BaseQuantizer
from qdiff.base.quant_layer import QuantizedLinear
from qdiff.base.quant_model import (bitwidth_refactor_, load_quant_param_dict_, quant_layer_refactor_,
                                    save_quant_param_dict_, set_init_done_)
from qdiff.utils import apply_func_to_submodules

logger = logging.getLogger(__name__)


class QuantWanMixin:
    quant_config = None

    # ---- quant_wanx.py:80-135 -----------------------------------------------------------------------------------
    def convert_quant(self, quant_config):
        self.quant_config = quant_config
        self.quant_param_dict = {}
        self.quant_layer_refactor()

    def quant_layer_refactor(self):
        apply_func_to_submodules(self, class_type=nn.Linear, function=quant_layer_refactor_, name=None, parent_module=None,
                                 quant_config=self.quant_config, full_name=None,
                                 remain_fp_regex=self.quant_config.remain_fp_regex)

    def save_quant_param_dict(self):
        apply_func_to_submodules(self, class_type=BaseQuantizer, function=save_quant_param_dict_, full_name=None,
                                 parent_module=None, model=self)

    def load_quant_param_dict(self, quant_param_dict):
        apply_func_to_submodules(self, class_type=BaseQuantizer, function=load_quant_param_dict_, full_name=None,
                                 parent_module=None, quant_param_dict=quant_param_dict, model=self)

    def set_init_done(self):
        apply_func_to_submodules(self, class_type=BaseQuantizer, function=set_init_done_)

    def bitwidth_refactor(self):
        apply_func_to_submodules(self, class_type=QuantizedLinear, function=bitwidth_refactor_, name=None,
                                 parent_module=None, quant_config=self.quant_config, full_name=None)

    # ---- quant_wanx.py:137-185 ----------------------------------------------------------------------------------
    def quantize_and_save_weight(self, save_path):
        from . import int_checkpoint
        for param in self.parameters():
            param.requires_grad_(False)
        sd = int_checkpoint.save_int_checkpoint(self, save_path)
        logger.info("Finished Saving the Quantized Checkpoint: %d entries -> %s", len(sd), save_path)
        return sd

    # ---- quant_wanx.py:188-228 ----------------------------------------------------------------------------------
    def wan_config(self):
        """Hyper-parameters of the wrapped model, from its attributes (WanModel keeps them, model.py:484-499) or shapes."""
        from .model import WanConfig
        blk = self.blocks[0]
        dim = getattr(self, "dim", None) or blk.self_attn.q.in_features
        get = lambda name, default: getattr(self, name, default)
        return WanConfig(dim=dim, ffn_dim=get("ffn_dim", blk.ffn[0].out_features),
                         num_heads=get("num_heads", dim // 128), num_layers=len(self.blocks),
                         in_dim=get("in_dim", 16), out_dim=get("out_dim", 16),
                         text_dim=get("text_dim", self.text_embedding[0].in_features), text_len=get("text_len", 512),
                         freq_dim=get("freq_dim", self.time_embedding[0].in_features),
                         patch_size=tuple(get("patch_size", (1, 2, 2))), eps=get("eps", 1e-6))

    def hardware_forward_refactor(self, load_path, seq_len=None, attn_quant=False, use_graph=True, sp=None):
        """Load the integer checkpoint and route forward() through the B200 integer runtime.  `seq_len` (the reference
        sizes its shared QuantParams scratch with it, quant_wanx.py:200) is accepted and unused: nothing is pre-sized."""
        from . import int_checkpoint
        from .model import GraphedDiT
        cfg = self.wan_config()
        sd = torch.load(load_path, map_location="cpu", weights_only=False) if isinstance(load_path, str) else load_path
        dit = int_checkpoint.dit_from_int_state_dict(cfg, sd, sp=sp, attn_quant=attn_quant)
        self._b200_dit = dit
        self._b200_step = GraphedDiT(dit) if use_graph else dit.forward
        self.forward = self._hardware_forward
        logger.info("hardware_forward_refactor: %d blocks on the integer runtime (%s attention)", cfg.num_layers,
                    "int8" if attn_quant else "bf16")
        return self

    def _hardware_forward(self, x, t, context, seq_len=None, clip_fea=None, y=None):
        """WanModel.forward's convention (wan/modules/model.py:539-631): x list of [C,F,H,W], t [B], context list of
        [T, text_dim] -> list of [C,F,H,W] fp32."""
        if clip_fea is not None or y is not None:
            raise NotImplementedError("the integer runtime covers the t2v path (no image conditioning)")
        outs = []
        for i, (u, c) in enumerate(zip(x, context)):
            o = self._b200_step(u.float(), t[i:i + 1].float(), c)
            outs.append(o.clone() if len(x) > 1 else o)       # the graph's output buffer is reused by the next replay
        return outs


from .wan_model import WanModelFP  # noqa: E402


class QuantWanModel(QuantWanMixin, WanModelFP):
    """quant_wanx.py:28-78: WanModel's constructor arguments plus `quant_config`; nothing is quantized until
    `quant_layer_refactor()` / `convert_quant(cfg)` is called (the reference leaves that call to the caller as well, :75-76)."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=1536, ffn_dim=8960, freq_dim=256,
                 text_dim=4096, out_dim=16, num_heads=12, num_layers=30, window_size=(-1, -1), qk_norm=True,
                 cross_attn_norm=True, eps=1e-6, quant_config=None):
        super().__init__(model_type=model_type, patch_size=patch_size, text_len=text_len, in_dim=in_dim, dim=dim,
                         ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim, num_heads=num_heads,
                         num_layers=num_layers, window_size=window_size, qk_norm=qk_norm, cross_attn_norm=cross_attn_norm, eps=eps)
        self.quant_config = quant_config
        self.quant_param_dict = {}
