"""List-valued bit-widths (e.g. weight.n_bits: [4, 8]) with per-bit-width parameters selected by
`bitwidth_refactor(i)` — mirror of ViDiT-Q/quant_utils/qdiff/base/mixed_precision_quantizer.py.
The reference leaves `n_bits`/`n_levels` stale after init / refactor; here
they always describe the selected bit-width (results are identical: its clamp never binds)."""
import logging

import torch

import b200q
from qdiff.base.base_quantizer import BaseQuantizer, _levels, _on_cuda

try:
    from omegaconf import ListConfig
except ImportError:  # pragma: no cover
    ListConfig = ()

logger = logging.getLogger(__name__)


class MixedPrecisionBaseQuantizer(BaseQuantizer):
    def __init__(self, quant_config):
        torch.nn.Module.__init__(self)
        bits = quant_config["n_bits"]
        assert isinstance(bits, (list, tuple)) or (ListConfig and isinstance(bits, ListConfig)), \
            "mixed-precision quantizers take a list of bit-widths"
        assert quant_config.get("i_bitwidth", None) is not None, "i_bitwidth selects the active entry of n_bits"
        self.bitwidth_list = bits
        self.i_bitwidth = quant_config["i_bitwidth"]
        self.n_bits = self.bitwidth_list[self.i_bitwidth]
        self.sym = quant_config.get("sym", False)
        self.register_buffer("delta_list", None)
        self.register_buffer("zero_point_list", None)
        self.register_buffer("delta", None)
        self.register_buffer("zero_point", None)
        self.n_levels = _levels(self.n_bits, self.sym)
        self.init_done = False
        self.module_name = None

    def bitwidth_refactor(self, i_bitwidth):
        # mixed_precision_quantizer.py:50-54
        self.i_bitwidth = i_bitwidth
        self.n_bits = self.bitwidth_list[i_bitwidth]
        self.n_levels = _levels(self.n_bits, self.sym)
        if self.delta_list is not None:
            self.delta = self.delta_list[i_bitwidth]
            self.zero_point = self.zero_point_list[i_bitwidth]


class MixedPrecisionStaticQuantizer(MixedPrecisionBaseQuantizer):
    """mixed_precision_quantizer.py:56-125"""

    def init_quant_params(self, x):
        assert x.dim() == 2
        xc, home = _on_cuda(x.detach())
        deltas, zps = [], []
        for bits in self.bitwidth_list:
            _, d, z, _ = b200q.quant_rows(xc, bits, self.sym, False, want_rowsum=False)
            deltas.append(d.unsqueeze(-1))
            zps.append(z.unsqueeze(-1))
        dl, zl = torch.stack(deltas, 0).to(x.dtype), torch.stack(zps, 0).to(x.dtype)
        if home is not None:
            dl, zl = dl.to(home), zl.to(home)
        self.delta_list, self.zero_point_list = dl, zl
        self.bitwidth_refactor(self.i_bitwidth)

    def quantize_int8(self, x):
        if self.init_done is not True:
            self.init_quant_params(x)
        xc, _ = _on_cuda(x.detach())
        return b200q.quant_rows_static(xc, self.delta.to(xc.device), self.zero_point.to(xc.device), self.n_bits, self.sym)[0]

    def quantize(self, x):
        _, home = _on_cuda(x.detach())
        return self._codes_as(self.quantize_int8(x), x, home)

    def forward(self, x):
        _, home = _on_cuda(x.detach())
        return self._dequant(self.quantize_int8(x), x, home)


class MixedPrecisionDynamicQuantizer(MixedPrecisionBaseQuantizer):
    """mixed_precision_quantizer.py:127-186 — dynamic: no parameter lists, only the active bit-width."""

    def quantize_int8(self, x, want_rowsum=True):
        assert x.dim() == 2
        return b200q.quant_rows(x, self.n_bits, self.sym, True, want_rowsum=want_rowsum)

    def quantize(self, x):
        q, _, _, _, home = self._compute(x, dynamic=True)
        return self._codes_as(q, x, home)

    def forward(self, x):
        q, _, _, _, home = self._compute(x, dynamic=True)
        return self._dequant(q, x, home)

    def bitwidth_refactor(self, i_bitwidth):
        self.i_bitwidth = i_bitwidth
        self.n_bits = self.bitwidth_list[i_bitwidth]
        self.n_levels = _levels(self.n_bits, self.sym)
