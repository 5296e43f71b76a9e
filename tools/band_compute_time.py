"""How long does ONE rank's share of the sharded 4K step take without any exchange?  Times the
unsharded engine on band-sized images (272 x 3840 = one of eight bands; 1080 x 3840 = one of two) on
a single GPU: the difference to the measured sharded step is what halo exchange, rank skew and the
Gram all-reduce cost.

    python tools/band_compute_time.py
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402


def main() -> None:
    from style_transfer_visualizer_b200.core_model import initialize_input
    from style_transfer_visualizer_b200.fused_step import FusedStep
    from style_transfer_visualizer_b200.optim import FusedAdam

    dev = torch.device("cuda:0")
    for (h, w) in [(272, 3840), (544, 3840), (1080, 3840), (2160, 3840)]:
        model, content_h, style_h = bench.build_job(dev, h, w, 0)
        content = content_h.to(dev)
        model.set_targets(style_h.to(dev), content)
        x = initialize_input(content, "content")
        fused = FusedStep.try_create(model, x, FusedAdam([x], lr=0.01), 1e5, 1.0)
        for _ in range(5):
            fused.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 30
        for _ in range(n):
            fused.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{h}x{w}: {ms:.3f} ms/step = {1e3 / ms:.1f} steps/s  ({fused.kernel_launches} launches)",
              flush=True)
        del fused, model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
