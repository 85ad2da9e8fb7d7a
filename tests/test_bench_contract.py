"""bench.py output contract (CPU part): the reference arm runs the oracle port of the reference's fake-quant block on the
host cores and prints ONE JSON line with the keys the driver reads; the CUDA arm refuses to run without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0",
                        "--ref-tokens", "1024"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False and d["value"] > 0
    ref_here = os.path.isdir("/root/reference/ViDiT-Q/quant_utils/qdiff")
    assert d["cpu_baseline"]["kind"] == ("reference-import" if ref_here else "port")
    assert d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["sample"]
    assert d["steps"] == 2 and len(d["block_ms_all"]) == 2 and d["value"] == pytest.approx(30 * d["block_ms"])
    assert "invalid" in d                                   # reduced-token CI run is marked as such
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU present")
def test_cuda_arm_has_no_cpu_fallback():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
