"""Steps/s of the reference's DEFAULT optimiser (L-BFGS, max_iter=1, history 100) through the
whole-step CUDA graph (FusedStep + FusedLBFGS.device_step), 512x512 and 1080p."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def main() -> None:
    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import synthetic
    from style_transfer_visualizer_b200.fused_step import FusedStep
    from style_transfer_visualizer_b200.optim import FusedLBFGS

    dev = torch.device("cuda:0")
    for (h, w, reps) in ((512, 512, 150), (1080, 1920, 40)):
        original = cm.initialize_vgg
        cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
        try:
            model = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
        finally:
            cm.initialize_vgg = original
        content = synthetic.synthetic_image(1, h, w).to(dev)
        style = synthetic.synthetic_image(2, h, w).to(dev)
        model.set_targets(style, content)
        x = cm.initialize_input(content, "random")
        opt = FusedLBFGS([x], lr=1.0, max_iter=1, max_eval=1)
        fs = FusedStep(model, x, opt, 1e5, 1.0)
        for _ in range(10):
            fs.step()
        torch.cuda.synchronize()
        for blk in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fs.step()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            print(f"L-BFGS graph step {h}x{w} block {blk}: {ms * 1e3:8.1f} us/step ({1e3 / ms:6.1f} steps/s) "
                  f"total loss {float(fs.scores[2]):.4e}", flush=True)
        del fs, opt, x, model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
