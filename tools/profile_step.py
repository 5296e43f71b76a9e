"""Run a few EAGER optimisation steps (no CUDA graph) so that a profiler sees every kernel.
Usage: python tools/profile_step.py --size 1080p|512 [--steps 2]   (wrap with ncu on the GPU box)"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    h, w = (1080, 1920) if args.size == "1080p" else (int(args.size), int(args.size))
    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import synthetic
    from style_transfer_visualizer_b200.optim import FusedAdam

    dev = torch.device("cuda:0")
    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
    finally:
        cm.initialize_vgg = original
    content = synthetic.synthetic_image(1, h, w).to(dev)
    style = synthetic.synthetic_image(2, h, w).to(dev)
    model.set_targets(style, content)
    x = cm.initialize_input(content, "content")
    opt = FusedAdam([x], lr=0.01)

    def closure():
        opt.zero_grad()
        sl, cl = model(x)
        loss = 1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()
        loss.backward()
        return loss

    opt.step(closure)  # warm-up
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        opt.step(closure)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
