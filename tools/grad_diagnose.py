"""Full-resolution gradient check of the CUDA path against the CPU oracle (run on the GPU box).

Golden fixtures above 128x128 store strided samples; this tool recomputes the whole first-closure
gradient with the oracle on the box's host cores and reports where the CUDA gradient deviates:
overall, by pixel parity class (pool windows), on the image border, and the worst pixel.

    python tools/grad_diagnose.py [H W] [--compact 0|1]
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import stv_oracle as orc  # noqa: E402


def main() -> None:
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    h, w = (int(args[0]), int(args[1])) if len(args) >= 2 else (1080, 1920)
    init = "content"
    if "--random" in sys.argv:
        init = "random"
    if "--compact" in sys.argv:
        import style_transfer_visualizer_b200.engine as eng

        eng.DEFAULT_COMPACT_BACKWARD = bool(int(sys.argv[sys.argv.index("--compact") + 1]))
    import style_transfer_visualizer_b200.core_model as cm

    dev = torch.device("cuda:0")
    content = orc.synthetic_image(1, h, w)
    style = orc.synthetic_image(2, h, w)
    x0 = content.clone() if init == "content" else \
        torch.randn(content.shape, generator=torch.Generator().manual_seed(3))
    torch.set_num_threads(torch.get_num_threads())
    t0 = time.time()
    om = orc.OracleModel(orc.vgg19_features(0))
    om.set_targets(style, content)
    ox = x0.clone().requires_grad_(True)
    s, c, t, sl, cl = orc.closure_step(om, ox, 1e5, 1.0)
    og = ox.grad.detach().clone()
    print(f"oracle: {time.time() - t0:.1f} s, total {float(t):.6e}, style layers "
          f"{[f'{float(v):.4e}' for v in sl]} content {[f'{float(v):.4e}' for v in cl]}", flush=True)

    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: orc.vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
    finally:
        cm.initialize_vgg = original
    model.set_targets(style.to(dev), content.to(dev))
    x = x0.to(dev).requires_grad_(True)
    gsl, gcl = model(x)
    (1e5 * torch.stack(gsl).sum() + torch.stack(gcl).sum()).backward()
    g = x.grad.detach().cpu()
    print(f"cuda  : style layers {[f'{float(v):.4e}' for v in gsl]} content "
          f"{[f'{float(v):.4e}' for v in gcl]}")
    d = (g - og).double()
    ref = og.double()

    def rel(mask=None) -> float:  # noqa: ANN001
        if mask is None:
            return float(d.norm() / ref.norm())
        return float(d[mask].norm() / ref[mask].norm())

    print(f"rel L2 full {rel():.4e}   |g| l2 cuda {float(g.double().norm()):.6e} oracle "
          f"{float(ref.norm()):.6e}   sum cuda {float(g.double().sum()):.6e} oracle "
          f"{float(ref.sum()):.6e}")
    yy = torch.arange(h).view(1, 1, h, 1).expand_as(og)
    xx = torch.arange(w).view(1, 1, 1, w).expand_as(og)
    for py in (0, 1):
        for px in (0, 1):
            print(f"  parity (y%2={py}, x%2={px}): rel L2 {rel((yy % 2 == py) & (xx % 2 == px)):.4e}")
    for k in (16, 8, 4, 2):
        for ry in range(0, k):
            m = (yy % k == ry)
            print(f"  rows y%{k}=={ry}: {rel(m):.3e}", end="")
        print()
    for name, m in (("top 4 rows", yy < 4), ("bottom 4 rows", yy >= h - 4), ("left 4 cols", xx < 4),
                    ("right 4 cols", xx >= w - 4), ("bottom 24 rows", yy >= h - 24),
                    ("interior", (yy >= 32) & (yy < h - 32) & (xx >= 32) & (xx < w - 32))):
        print(f"  {name}: rel L2 {rel(m):.4e}")
    idx = int(d.abs().argmax())
    cch, yq, xq = np.unravel_index(idx, tuple(og.shape[1:]))
    print(f"  worst pixel: c={cch} y={yq} x={xq} cuda {float(g.flatten()[idx]):.4e} oracle "
          f"{float(og.flatten()[idx]):.4e}  (max |g| oracle {float(og.abs().max()):.4e})")
    # error energy per row band of 8 rows (top 5)
    rows = d.pow(2).sum(dim=(0, 1, 3))
    rref = ref.pow(2).sum(dim=(0, 1, 3))
    ratio = (rows / rref).sqrt()
    top = torch.topk(ratio, 8)
    print("  worst rows (rel):", [(int(i), f"{float(v):.2e}") for v, i in zip(top.values, top.indices)])
    cols = d.pow(2).sum(dim=(0, 1, 2))
    cref = ref.pow(2).sum(dim=(0, 1, 2))
    top = torch.topk((cols / cref).sqrt(), 8)
    print("  worst cols (rel):", [(int(i), f"{float(v):.2e}") for v, i in zip(top.values, top.indices)])


if __name__ == "__main__":
    main()
