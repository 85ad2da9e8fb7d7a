"""Import the UNMODIFIED reference `qdiff` package from /root/reference (exists only in
the build container, never on the GPU box).  Used by oracle/gen_golden.py and by
tests/test_oracle_vs_reference.py to pin the oracle restatement.

Recipe: the reference imports `omegaconf.ListConfig` at module top
(quant_utils/qdiff/base/base_quantizer.py:9); when omegaconf is not installed we
put the in-repo shim on sys.path first.  The reference package is exposed under the
module name `qdiff`, so it must not be mixed in one interpreter with the product's
own `qdiff` mirror: callers run it in a subprocess or before importing the mirror.
"""
import importlib
import os
import sys

REF_ROOT = "/root/reference/ViDiT-Q/quant_utils"
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(os.path.dirname(_HERE), "wan2.1-quantization_b200", "compat")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "qdiff"))


def import_reference_qdiff():
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    try:
        importlib.import_module("omegaconf")
    except ImportError:
        sys.path.insert(0, _SHIM)
    for name in list(sys.modules):
        if name == "qdiff" or name.startswith("qdiff."):
            mod = sys.modules[name]
            if REF_ROOT not in (getattr(mod, "__file__", "") or ""):
                raise RuntimeError("a non-reference `qdiff` is already imported in this interpreter")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import qdiff.base.base_quantizer as bq
    import qdiff.base.quant_layer as ql
    import qdiff.base.mixed_precision_quantizer as mp
    import qdiff.base.quant_attn as qa
    import qdiff.utils as qu
    return dict(base_quantizer=bq, quant_layer=ql, mixed_precision=mp, quant_attn=qa, utils=qu)
