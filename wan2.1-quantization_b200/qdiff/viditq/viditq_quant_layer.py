"""ViDiT-Q layer = smooth scale then rotate (ViDiT-Q/quant_utils/qdiff/viditq/viditq_quant_layer.py:30-73) —
SURVEY §8 (f)-2, a "next" row.  The shipped YAML selects it for every layer (quant_configs/config.yaml:19-21);
until it is built, drop the `viditq` section (the reference's own hardware path also ignores the mask and
the rotation: SURVEY appendix B-4)."""
from qdiff.base.quant_layer import QuantizedLinear


class ViDiTQuantizedLinear(QuantizedLinear):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "viditq layers are a next-row item (SURVEY §8f-2) not built yet: remove the `viditq` section from the "
            "quant_config to run plain W8A8/W4A8 QuantizedLinear layers")
