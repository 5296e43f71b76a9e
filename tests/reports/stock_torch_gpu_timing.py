"""The "kernel to beat" figure of SURVEY section 8(d): the reference's own algorithm (oracle port:
torchvision VGG19 modules, torch.mm Gram, autograd, torch.optim.Adam) timed on cuda:0 through STOCK
PyTorch -- cuDNN TF32 convolutions, fp32 cuBLAS mm, eager launches, the reference runner's closure
shape -- next to this package on the same inputs.  Test infrastructure (uses oracle/); not a bench
value.

    python tests/reports/stock_torch_gpu_timing.py [512 1080p]
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from oracle import stv_oracle as orc  # noqa: E402


def time_stock(h: int, w: int, steps: int, dev: torch.device) -> float:
    feats = orc.vgg19_features(0).to(dev)
    model = orc.OracleModel(feats, [0, 5, 10, 19, 28], [21])
    content = orc.synthetic_image(1, h, w).to(dev)
    style = orc.synthetic_image(2, h, w).to(dev)
    model.set_targets(style, content)
    x = content.clone().requires_grad_(True)
    opt = torch.optim.Adam([x], lr=0.01)

    def closure():  # noqa: ANN202
        opt.zero_grad()
        sl, cl = model(x)
        loss = 1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()
        loss.backward()
        return loss

    for _ in range(3):
        opt.step(closure)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        opt.step(closure)
    torch.cuda.synchronize(dev)
    return steps / (time.perf_counter() - t0)


def time_ours(h: int, w: int, steps: int, dev: torch.device) -> float:
    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200.fused_step import FusedStep
    from style_transfer_visualizer_b200.optim import FusedAdam

    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: orc.vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
    finally:
        cm.initialize_vgg = original
    content = orc.synthetic_image(1, h, w).to(dev)
    style = orc.synthetic_image(2, h, w).to(dev)
    model.set_targets(style, content)
    x = cm.initialize_input(content, "content")
    fused = FusedStep.try_create(model, x, FusedAdam([x], lr=0.01), 1e5, 1.0)
    for _ in range(3):
        fused.step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        fused.step()
    torch.cuda.synchronize(dev)
    return steps / (time.perf_counter() - t0)


def main() -> None:
    dev = torch.device("cuda:0")
    sizes = sys.argv[1:] or ["512", "1080p"]
    print(f"torch {torch.__version__}, cudnn {torch.backends.cudnn.version()}, "
          f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32} "
          f"matmul.allow_tf32={torch.backends.cuda.matmul.allow_tf32}, "
          f"{torch.cuda.get_device_name(0)}")
    for s in sizes:
        h, w = (1080, 1920) if s == "1080p" else (int(s), int(s))
        steps = 30 if h * w > 1_000_000 else 100
        torch.backends.cudnn.benchmark = False
        a = time_stock(h, w, steps, dev)
        torch.backends.cudnn.benchmark = True
        b = time_stock(h, w, steps, dev)
        c = time_ours(h, w, steps, dev)
        print(f"{h}x{w}: stock PyTorch {a:7.1f} steps/s (cudnn.benchmark=True: {b:7.1f}) | "
              f"this package {c:7.1f} steps/s  ({c / max(a, b):.2f}x)", flush=True)


if __name__ == "__main__":
    main()
