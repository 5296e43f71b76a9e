"""Kernel-level parity on a real B200: every C-ABI entry point against plain PyTorch fp32.

The cases live in tools/gpu_selftest.py (also usable stand-alone under gpurun); each runs in its
own subprocess so a trapping kernel cannot poison the remaining tests' CUDA context.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["elementwise", "conv", "conv_variants", "gram"])
def test_kernel_case(case: str) -> None:
    proc = subprocess.run([sys.executable, str(ROOT / "tools" / "gpu_selftest.py"), "--case", case],
                          capture_output=True, text=True, timeout=600, check=False, cwd=ROOT)
    fails = [ln for ln in proc.stdout.splitlines() if ln.startswith("FAIL")]
    assert proc.returncode == 0 and not fails, \
        "\n".join(fails) + "\n" + proc.stdout[-3000:] + proc.stderr[-3000:]
    assert proc.stdout.count("PASS") > 5
