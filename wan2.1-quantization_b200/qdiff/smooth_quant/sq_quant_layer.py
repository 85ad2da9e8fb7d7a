"""SmoothQuant layer (ViDiT-Q/quant_utils/qdiff/smooth_quant/sq_quant_layer.py) — SURVEY §8 (f)-2, a "next" row:
per-channel mask = |W|max^alpha / |X|max^(1-alpha) (:27-34), W/mask re-quantised (:36-44), X*mask before the
activation quantizer (:55).  Not built in this round; selecting it fails loudly instead of silently running
the plain layer."""
from qdiff.base.quant_layer import QuantizedLinear


class SQQuantizedLinear(QuantizedLinear):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "smooth_quant layers are a next-row item (SURVEY §8f-2) not built yet: remove the `smooth_quant` section "
            "from the quant_config to run plain W8A8/W4A8 QuantizedLinear layers")
