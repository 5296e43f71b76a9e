"""Time the default-optimiser path (FusedLBFGS, reference defaults) vs torch.optim.LBFGS on the
same model: steps/s as the history fills (run on the GPU box)."""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import style_transfer_visualizer_b200.core_model as cm  # noqa: E402
from style_transfer_visualizer_b200 import synthetic  # noqa: E402
from style_transfer_visualizer_b200.optim import FusedLBFGS  # noqa: E402


def main() -> None:
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    dev = torch.device("cuda:0")
    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
    finally:
        cm.initialize_vgg = original
    content = synthetic.synthetic_image(1, size, size).to(dev)
    style = synthetic.synthetic_image(2, size, size).to(dev)
    model.set_targets(style, content)
    from style_transfer_visualizer_b200.config import StyleTransferConfig
    from style_transfer_visualizer_b200.optimization import OptimizationRunner

    class Bar:
        def update(self, n=1): ...
        def set_postfix(self, *a, **k): ...
        def close(self): ...

    torch.manual_seed(0)
    xg = torch.randn_like(content).requires_grad_(True)
    cfg = StyleTransferConfig.model_validate({
        "optimization": {"steps": steps, "style_w": 1e9, "content_w": 1.0, "lr": 1.0},
        "video": {"save_every": steps + 1}, "output": {"log_every": 50}})
    marks = []
    runner = OptimizationRunner(model, xg, cfg, progress_bar=Bar(), use_cuda_graph=True)
    from style_transfer_visualizer_b200.optimization import OptimizationCallbacks
    t0 = [0.0]

    def on_end(m):
        if m.step % 50 == 0:
            torch.cuda.synchronize()
            marks.append((m.step, time.perf_counter() - t0[0], m.total_loss))

    runner.callbacks = OptimizationCallbacks(on_step_end=on_end)
    runner.prepare()
    torch.cuda.synchronize()
    t0[0] = time.perf_counter()
    runner.run()
    prev = 0.0
    for i, t, l in marks:
        print(f"graph LBFGS {size}x{size} (runner default optimiser): steps {i - 49}-{i}: "
              f"{50 / (t - prev):.1f} steps/s  loss {l:.4e}", flush=True)
        prev = t
    print("counters", runner.optimizer.device_counters(), flush=True)

    for kind in ("fused", "torch"):
        torch.manual_seed(0)
        x = torch.randn_like(content).requires_grad_(True)
        opt = FusedLBFGS([x], lr=1.0, max_iter=1, max_eval=1) if kind == "fused" else \
            torch.optim.LBFGS([x], lr=1.0, max_iter=1, max_eval=1)

        def closure():
            opt.zero_grad()
            sl, cl = model(x)
            loss = 1e9 * torch.stack(sl).sum() + torch.stack(cl).sum()
            loss.backward()
            return loss

        marks = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(1, steps + 1):
            loss = opt.step(closure)
            if i % 50 == 0:
                torch.cuda.synchronize()
                marks.append((i, time.perf_counter() - t0, float(loss)))
        prev = 0.0
        for i, t, l in marks:
            print(f"{kind} LBFGS {size}x{size}: steps {i - 49}-{i}: {50 / (t - prev):.1f} steps/s  loss {l:.4e}",
                  flush=True)
            prev = t


if __name__ == "__main__":
    main()
