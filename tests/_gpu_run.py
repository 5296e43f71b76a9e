"""Run a golden case through the CUDA implementation (shared by GPU tests and report tools)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from oracle import stv_oracle as orc
from tests import _cases as cases


class MemorySink:
    """In-memory VideoFrameSink."""

    def __init__(self) -> None:
        self.frames: list[np.ndarray] = []
        self._size = None
        self.closed = False

    def append_data(self, frame: np.ndarray) -> None:
        self.frames.append(np.array(frame, copy=True))

    def close(self) -> None:
        self.closed = True


class NullBar:
    def update(self, n=1): ...  # noqa: ANN001, ANN201
    def set_postfix(self, *a, **k): ...  # noqa: ANN002, ANN003, ANN201
    def close(self): ...  # noqa: ANN201


@dataclass
class GpuResult:
    layer_style: list[float] = field(default_factory=list)
    layer_content: list[float] = field(default_factory=list)
    first_grad: np.ndarray | None = None
    style: list[float] = field(default_factory=list)
    content: list[float] = field(default_factory=list)
    total: list[float] = field(default_factory=list)
    final: np.ndarray | None = None
    frames: list[np.ndarray] = field(default_factory=list)
    csv_text: str = ""


def build_model(cfg: dict, device: torch.device):  # noqa: ANN201
    """This package's StyleContentModel with the seeded random-init VGG19 (as the reference's
    tests patch ``initialize_vgg``) and targets set from the case's images."""
    import style_transfer_visualizer_b200.core_model as cm

    content, style, init = cases.case_inputs(cfg)
    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: orc.vgg19_features(cfg["weight_seed"])
    try:
        model = cm.StyleContentModel(cfg["style_layers"], cfg["content_layers"]).to(device)
    finally:
        cm.initialize_vgg = original
    model.set_targets(style.to(device), content.to(device))
    x0 = cases.initial_image(cfg, content, init).to(device)
    return model, x0


def run_case(cfg: dict, device: torch.device, *, use_cuda_graph: bool | None = False,
             csv_path: str | None = None) -> GpuResult:
    from style_transfer_visualizer_b200.config import StyleTransferConfig
    from style_transfer_visualizer_b200.optim import FusedAdam, FusedLBFGS
    from style_transfer_visualizer_b200.optimization import OptimizationRunner

    res = GpuResult()
    model, x0 = build_model(cfg, device)
    style_w = cfg.get("style_w", 1e5)
    content_w = cfg.get("content_w", 1.0)

    probe = x0.clone().requires_grad_(True)
    sl, cl = model(probe)
    (style_w * torch.stack(sl).sum() + content_w * torch.stack(cl).sum()).backward()
    res.layer_style = [float(v.detach()) for v in sl]
    res.layer_content = [float(v.detach()) for v in cl]
    res.first_grad = probe.grad.detach().cpu().numpy()

    x = x0.clone().requires_grad_(True)
    if cfg["opt"] == "adam":
        opt = FusedAdam([x], lr=cfg["lr"])
    else:
        opt = FusedLBFGS([x], lr=cfg["lr"], max_iter=1, max_eval=1)
    config = StyleTransferConfig.model_validate({
        "optimization": {"steps": cfg["steps"], "style_w": style_w, "content_w": content_w,
                         "lr": cfg["lr"], "init_method": cfg["init"],
                         "normalize": cfg.get("normalize", True),
                         "style_layers": cfg["style_layers"],
                         "content_layers": cfg["content_layers"]},
        "video": {"save_every": cfg.get("save_every") or cfg["steps"] + 1},
        "output": {"log_every": 1, "log_loss": csv_path},
    })
    sink = MemorySink() if cfg.get("save_every") else None
    runner = OptimizationRunner(model, x, config, optimizer=opt, progress_bar=NullBar(),
                                video_writer=sink, use_cuda_graph=use_cuda_graph)
    final, history, _elapsed = runner.run()
    assert final is x
    if csv_path:
        from pathlib import Path

        res.csv_text = Path(csv_path).read_text(encoding="utf-8")
        rows = [r.split(",") for r in res.csv_text.strip().splitlines()[1:]]
        history = {"style_loss": [float(r[1]) for r in rows],
                   "content_loss": [float(r[2]) for r in rows],
                   "total_loss": [float(r[3]) for r in rows]}
    res.style = history["style_loss"]
    res.content = history["content_loss"]
    res.total = history["total_loss"]
    res.final = final.detach().cpu().numpy()
    if sink is not None:
        res.frames = sink.frames
    return res


def compare(cfg: dict, gold: dict, res: GpuResult) -> dict[str, float]:
    """Error metrics of a GPU run against the reference's golden outputs."""
    out: dict[str, float] = {}
    ls, gs = np.array(res.layer_style), gold["layer_style"]
    out["layer_style_rel_max"] = float(np.max(np.abs(ls - gs) / (np.abs(gs) + 1e-300))) \
        if gs.size else 0.0
    lc, gc = np.array(res.layer_content), gold["layer_content"]
    scale = max(float(np.abs(gold["total_loss"][0])), 1e-300)
    out["layer_content_abs_over_total"] = float(np.max(np.abs(lc - gc))) / scale if gc.size else 0.0
    g = cases.subsample_like_golden(cfg, res.first_grad)
    out["grad_rel_l2"] = cases.rel_l2(g, gold["first_grad"])
    out["grad_cosine"] = cases.cosine(g, gold["first_grad"])
    tot, gtot = np.array(res.total), gold["total_loss"]
    out["total_rel_max"] = float(np.max(np.abs(tot - gtot) / (np.abs(gtot) + 1e-300)))
    sty, gsty = np.array(res.style), gold["style_loss"]
    out["style_rel_max"] = float(np.max(np.abs(sty - gsty) / (np.abs(gsty) + 1e-300)))
    f = cases.subsample_like_golden(cfg, res.final)
    out["final_rel_l2"] = cases.rel_l2(f, gold["final"])
    out["final_max_abs"] = float(np.max(np.abs(f - gold["final"])))
    # movement-relative error: |x_mine - x_ref| / |x_ref - x_0|
    if "frames" in gold and gold["frames"].size and res.frames:
        fr = np.stack(res.frames)
        k = cases.sample_stride(cfg)
        if k > 1:
            fr = fr[:, ::k, ::k, :]
        d = np.abs(fr.astype(int) - gold["frames"].astype(int))
        out["frames_max_lsb"] = float(d.max())
        out["frames_frac_diff"] = float((d > 0).mean())
    return out
