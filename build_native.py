"""Build libstv_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "style_transfer_visualizer_b200" / "csrc"
OUT = ROOT / "style_transfer_visualizer_b200" / "lib" / "libstv_b200.so"
SOURCES = ["api.cu", "conv_igemm2.cu", "gram.cu", "conv_direct.cu", "conv_first_tc.cu", "conv_first_dgrad_tc.cu", "elementwise.cu",
           "lbfgs.cu", "halo.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_rebuild() -> bool:
    if not OUT.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.iterdir()) + [ROOT / "include" / "stv_b200.h"])
    return newest > OUT.stat().st_mtime


def build(*, force: bool = False, verbose: bool = False, experiments: bool = False) -> Path:
    """``experiments=True`` compiles the bottleneck-experiment switches in (``-DSTV_EXPERIMENTS``:
    the conv kernel then honours the STV_CONV_DEBUG environment variable and produces garbage
    when it is set) -- for tools/conv_bottleneck.py and tools/conv_chain_latency.py only; the
    product build never reads the environment."""
    if not force and not experiments and not needs_rebuild():
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    import os
    import shlex

    extra = shlex.split(os.environ.get("STV_NVCC_EXTRA", ""))  # A/B builds of tools/ scripts only
    cmd = ["nvcc", *FLAGS, *extra, *(["-DSTV_EXPERIMENTS"] if experiments else []),
           *(["-Xptxas", "-v"] if verbose else []), "-o", str(OUT),
           *[str(CSRC / s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True, check=False)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        msg = f"nvcc failed with exit code {proc.returncode}"
        raise RuntimeError(msg)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                experiments="--experiments" in sys.argv))
