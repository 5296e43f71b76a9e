"""Context for the stated TF32 tolerance: how far does the REFERENCE'S OWN algorithm move when it
runs on CUDA through stock PyTorch (cuDNN TF32 convolutions + fp32 cuBLAS mm, default flags) instead
of the CPU?  Prints, per golden case, the deviation from the CPU-fp32 fixtures of (a) the oracle
port executed on cuda:0 and (b) this package's kernels.  Test infrastructure (uses oracle/)."""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import stv_oracle as orc  # noqa: E402
from tests import _cases as cases  # noqa: E402
from tests import _gpu_run  # noqa: E402


def main() -> None:
    dev = torch.device("cuda:0")
    print(f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32} "
          f"matmul.allow_tf32={torch.backends.cuda.matmul.allow_tf32}")
    for name in cases.golden_names():
        cfg, gold = cases.load_golden(name)
        if cfg["opt"] != "adam":
            continue
        content, style, init = cases.case_inputs(cfg)
        feats = orc.vgg19_features(cfg["weight_seed"]).to(dev)
        model = orc.OracleModel(feats, cfg["style_layers"], cfg["content_layers"])
        model.set_targets(style.to(dev), content.to(dev))
        x = cases.initial_image(cfg, content, init).to(dev).requires_grad_(True)
        sl, cl = model(x)
        (cfg.get("style_w", 1e5) * torch.stack(sl).sum() + torch.stack(cl).sum()).backward()
        ls = np.array([float(v.detach()) for v in sl])
        g = cases.subsample_like_golden(cfg, x.grad)
        t_style = float(np.max(np.abs(ls - gold["layer_style"]) / np.abs(gold["layer_style"])))
        t_grad = cases.rel_l2(g, gold["first_grad"])
        res = _gpu_run.run_case(cfg, dev, use_cuda_graph=False)
        m = _gpu_run.compare(cfg, gold, res)
        print(f"{name:28s} torch-cuda(cuDNN tf32): style_rel_max={t_style:.2e} grad_rel_l2={t_grad:.2e}"
              f" | this package: style_rel_max={m['layer_style_rel_max']:.2e} "
              f"grad_rel_l2={m['grad_rel_l2']:.2e}", flush=True)


if __name__ == "__main__":
    main()
