"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b200q.h declares.
No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200q.h")


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"B200Q_API\s+(?:const\s+char\*|int)\s+(b200q_\w+)\s*\(", text)))


def test_header_declares_the_hot_path():
    names = _declared()
    for must in ("b200q_quant_rows", "b200q_calib_absmax_minmax", "b200q_gemm_w8a8", "b200q_gemm_w4a8", "b200q_pack_w4",
                 "b200q_ln_mod_quant", "b200q_gate_residual", "b200q_last_error", "b200q_version"):
        assert must in names


def test_library_builds_and_exports_every_declared_symbol():
    import b200q
    if not os.path.exists(b200q.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("b200q_build", os.path.join(ROOT, "wan2.1-quantization_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(b200q.LIB_PATH)
    declared = _declared()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200q.h but not exported"
    assert declared == b200q.exported_symbols(), "ctypes signature table and header disagree"
    assert b200q.version() >= (0, 1, 0)


def test_sass_is_blackwell_native():
    """tcgen05.mma.kind::i8 -> UTCIMMA, kind::f16 -> UTCHMMA, TMA -> UTMALDG/UTMASTG, tcgen05.ld -> LDTM (B200_PROFILING.md)."""
    import shutil
    import subprocess
    import b200q
    if shutil.which("cuobjdump") is None or not os.path.exists(b200q.LIB_PATH):
        pytest.skip("cuobjdump or library missing")
    sass = subprocess.run(["cuobjdump", "-sass", b200q.LIB_PATH], capture_output=True, text=True).stdout
    import re
    for mnemonic in ("UTCIMMA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "FFMA2"):
        assert mnemonic in sass, mnemonic
    # no legacy mma.sync path (UTCHMMA = tcgen05.mma.kind::f16 is the Blackwell one)
    assert re.search(r"(?<!UTC)HMMA", sass) is None and "IMMA.16" not in sass


def test_no_cpu_fallback():
    import b200q
    with pytest.raises(b200q.B200QError):
        b200q.quant_rows(torch.zeros(4, 16), 8, True, True)
    with pytest.raises(b200q.B200QError):
        b200q.gemm_w8a8(torch.zeros(4, 16, dtype=torch.int8), torch.zeros(4, 16, dtype=torch.int8), out_dtype=torch.int32)
