"""Print error metrics of the CUDA path against every golden fixture (run on the GPU box)."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from tests import _cases as cases  # noqa: E402
from tests import _gpu_run  # noqa: E402


def lbfgs_vs_torch(dev: torch.device) -> None:
    """FusedLBFGS against torch.optim.LBFGS, both driven by THIS package's model (same gradients):
    isolates the optimiser logic from TF32-vs-fp32 gradient differences."""
    from style_transfer_visualizer_b200.optim import FusedLBFGS

    cfg, _gold = cases.load_golden("lbfgs_random_64")
    model, x0 = _gpu_run.build_model(cfg, dev)
    hist = {}
    for kind in ("fused", "torch"):
        x = x0.clone().requires_grad_(True)
        opt = FusedLBFGS([x], lr=1.0, max_iter=1, max_eval=1) if kind == "fused" else \
            torch.optim.LBFGS([x], lr=1.0, max_iter=1, max_eval=1)
        losses = []

        def closure():
            opt.zero_grad()
            sl, cl = model(x)
            loss = cfg["style_w"] * torch.stack(sl).sum() + torch.stack(cl).sum()
            loss.backward()
            losses.append(float(loss.detach()))
            return loss

        for _ in range(12):
            opt.step(closure)
        hist[kind] = (losses, x.detach().clone())
    lf, lt = hist["fused"][0], hist["torch"][0]
    print("lbfgs fused :", " ".join(f"{v:.5e}" for v in lf))
    print("lbfgs torch :", " ".join(f"{v:.5e}" for v in lt))
    dx = (hist["fused"][1] - hist["torch"][1]).norm() / (hist["torch"][1] - x0).norm()
    print(f"lbfgs final image: |fused - torch| / |torch - x0| = {float(dx):.3e}", flush=True)


def main() -> None:
    dev = torch.device("cuda:0")
    if "--lbfgs" in sys.argv:
        sys.argv.remove("--lbfgs")
        lbfgs_vs_torch(dev)
    names = sys.argv[1:] or cases.golden_names()
    for name in names:
        cfg, gold = cases.load_golden(name)
        for graph in ((False, True) if cfg["opt"] == "adam" else (False,)):
            try:
                res = _gpu_run.run_case(cfg, dev, use_cuda_graph=graph)
                m = _gpu_run.compare(cfg, gold, res)
                print(f"{name} graph={graph}: " + " ".join(f"{k}={v:.3e}" for k, v in m.items()),
                      flush=True)
                if cfg["opt"] == "lbfgs":
                    print("    gpu:", " ".join(f"{v:.4e}" for v in res.total))
                    print("    ref:", " ".join(f"{v:.4e}" for v in gold["total_loss"]), flush=True)
                else:
                    print(f"    total gpu={res.total[0]:.6e}..{res.total[-1]:.6e} "
                          f"ref={gold['total_loss'][0]:.6e}..{gold['total_loss'][-1]:.6e}",
                          flush=True)
            except Exception as exc:  # noqa: BLE001
                import traceback

                traceback.print_exc()
                print(f"{name} graph={graph}: ERROR {type(exc).__name__}: {exc}", flush=True)


if __name__ == "__main__":
    main()
