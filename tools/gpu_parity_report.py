"""Print error metrics of the CUDA path against every golden fixture (run on the GPU box)."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from tests import _cases as cases  # noqa: E402
from tests import _gpu_run  # noqa: E402


def main() -> None:
    dev = torch.device("cuda:0")
    names = sys.argv[1:] or cases.golden_names()
    for name in names:
        cfg, gold = cases.load_golden(name)
        for graph in ((False, True) if cfg["opt"] == "adam" else (False,)):
            try:
                res = _gpu_run.run_case(cfg, dev, use_cuda_graph=graph)
                m = _gpu_run.compare(cfg, gold, res)
                print(f"{name} graph={graph}: " + " ".join(f"{k}={v:.3e}" for k, v in m.items()),
                      flush=True)
                print(f"    total gpu={res.total[0]:.6e}..{res.total[-1]:.6e} "
                      f"ref={gold['total_loss'][0]:.6e}..{gold['total_loss'][-1]:.6e}", flush=True)
            except Exception as exc:  # noqa: BLE001
                import traceback

                traceback.print_exc()
                print(f"{name} graph={graph}: ERROR {type(exc).__name__}: {exc}", flush=True)


if __name__ == "__main__":
    main()
