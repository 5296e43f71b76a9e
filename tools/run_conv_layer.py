"""Launch the hot-path conv forward for one layer shape a few times (for ncu captures)."""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from style_transfer_visualizer_b200 import ops  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="1080x1920x64x64,540x960x128x128,270x480x256x256")
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    for spec in args.shapes.split(","):
        h, w, cin, cout = (int(v) for v in spec.split("x"))
        x = torch.randn(h, w, cin, device=dev, generator=g)
        wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        b = torch.zeros(cout, device=dev)
        wf, _wd = ops.pack_conv_weights(wt)
        post = torch.empty(h, w, cout, device=dev)
        for _ in range(args.iters):
            ops.conv3x3_fwd(x, wf, b, None, post)
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
