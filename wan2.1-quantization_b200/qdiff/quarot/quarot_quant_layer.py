"""QuaRot layer (ViDiT-Q/quant_utils/qdiff/quarot/quarot_quant_layer.py) — SURVEY §8 (f)-2, a "next" row:
random-Hadamard rotation of W (fp64, :30-45) and of X per forward (:60).  Not built in this round."""
from qdiff.base.quant_layer import QuantizedLinear


class QuarotQuantizedLinear(QuantizedLinear):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "quarot layers are a next-row item (SURVEY §8f-2) not built yet: remove the `quarot` section from the "
            "quant_config to run plain W8A8/W4A8 QuantizedLinear layers")
