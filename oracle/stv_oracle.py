"""CPU oracle for the style-transfer optimisation step.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch fp32 on the CPU, the algorithm of the reference's hot path
(bjg-gh/style_transfer_visualizer, paths below relative to ``src/style_transfer_visualizer/``).
It exists so that the CUDA path can be checked on machines where ``/root/reference`` is absent
(the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; nothing under ``style_transfer_visualizer_b200/`` does.

Parity pin: the reference's own tests hold no numerical golden vectors for this path (SURVEY.md
section 8c), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in
the build container by ``oracle/make_golden.py`` (which imports the unmodified reference through
``oracle/reference_shim.py``) and committed under ``tests/golden/``.  ``tests/test_oracle.py``
checks this file against those fixtures (and against the live reference when it is present).

The arithmetic itself lives in third-party code the reference pins (torch 2.9.0, torchvision
0.24.0; this image has torch 2.11.0 / torchvision 0.26.0): Conv2d/ReLU/MaxPool2d, torch.mm,
mse_loss, autograd, torch.optim.Adam/LBFGS.  The restatement calls the same public torch ops at
the reference's call sites.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F  # noqa: N812
from torch import nn

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # constants.py:11
IMAGENET_STD = (0.229, 0.224, 0.225)   # constants.py:12
GRAM_MATRIX_CLAMP_MAX = 5e5            # constants.py:15
DEFAULT_STYLE_LAYERS = (0, 5, 10, 19, 28)  # config_defaults.py:18
DEFAULT_CONTENT_LAYERS = (21,)             # config_defaults.py:19

# torchvision vgg.py cfg "E" (VGG19); "M" = MaxPool2d(2, 2)
VGG19_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
             512, 512, 512, 512, "M")


def vgg19_features(seed: int) -> nn.Sequential:
    """Random-init, frozen, eval-mode ``vgg19().features`` (pretrained weights are unavailable
    offline).  Mirrors how tests/test_core_model.py:149-157 patches ``initialize_vgg``
    (core_model.py:103-117): same module list, ``requires_grad_(False)`` on every parameter."""
    torch.manual_seed(seed)
    try:
        from torchvision.models import vgg19  # same constructor the reference calls

        feats = vgg19(weights=None).features
    except ImportError:  # pragma: no cover - torchvision ships in this image
        layers: list[nn.Module] = []
        cin = 3
        for v in VGG19_CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                conv = nn.Conv2d(cin, int(v), kernel_size=3, padding=1)
                nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(conv.bias, 0)
                layers += [conv, nn.ReLU(inplace=True)]
                cin = int(v)
        feats = nn.Sequential(*layers)
    feats = feats.eval()
    for p in feats.parameters():
        p.requires_grad_(False)
    return feats


def synthetic_image(seed: int, height: int, width: int, *, normalize: bool = True,
                    scale: float = 1.0) -> torch.Tensor:
    """Uniform-random pixels passed through the reference's load transform
    (image_io.py:64-84: ToTensor then Normalize(mean, std))."""
    gen = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, height, width, generator=gen)
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        img = (img - mean) / std
    return img * scale


def gram_matrix(t: torch.Tensor, clamp_max: float = GRAM_MATRIX_CLAMP_MAX) -> torch.Tensor:
    """core_model.py:29-63 -- batch folded into channels, clamp BEFORE the 1/N."""
    b, c, h, w = t.shape
    flat = t.reshape(b * c, h * w)
    raw = torch.mm(flat, flat.t())
    return raw.clamp(max=clamp_max).div(b * c * h * w)


def split_blocks(features: nn.Module, style_layers, content_layers):  # noqa: ANN001, ANN201
    """core_model.py:120-146 -- cut the layer list after every tapped index; in-place ReLUs are
    replaced by out-of-place ones; layers after the last tap are dropped."""
    wanted = set(style_layers) | set(content_layers)
    blocks: list[nn.Sequential] = []
    style_ids: list[int] = []
    content_ids: list[int] = []
    pending: list[nn.Module] = []
    for idx, layer in enumerate(features.children()):
        pending.append(nn.ReLU(inplace=False) if isinstance(layer, nn.ReLU) else layer)
        if idx in wanted:
            blocks.append(nn.Sequential(*pending))
            pending = []
            if idx in style_layers:
                style_ids.append(len(blocks) - 1)
            if idx in content_layers:
                content_ids.append(len(blocks) - 1)
    return blocks, content_ids, style_ids


class OracleModel:
    """core_model.py:149-328 (StyleContentModel), restated without nn.Module plumbing."""

    def __init__(self, features: nn.Module, style_layers=DEFAULT_STYLE_LAYERS,  # noqa: ANN001
                 content_layers=DEFAULT_CONTENT_LAYERS) -> None:
        self.blocks, self.content_ids, self.style_ids = split_blocks(
            features, list(style_layers), list(content_layers))
        self.style_targets: list[torch.Tensor] | None = None
        self.content_targets: list[torch.Tensor] | None = None

    def set_targets(self, style_img: torch.Tensor, content_img: torch.Tensor) -> None:
        """core_model.py:192-232 -- two forwards; Grams of the style image, raw features of the
        content image, all detached."""
        grams = []
        x = style_img
        for j, blk in enumerate(self.blocks):
            x = blk(x)
            if j in self.style_ids:
                grams.append(gram_matrix(x).detach())
        feats = []
        x = content_img
        for j, blk in enumerate(self.blocks):
            x = blk(x)
            if j in self.content_ids:
                feats.append(x.detach())
        self.style_targets, self.content_targets = grams, feats

    def __call__(self, x: torch.Tensor):  # noqa: ANN204
        """core_model.py:297-328 -- returns (style_losses, content_losses), lists of 0-dim."""
        if self.style_targets is None or self.content_targets is None:
            msg = "targets must be set before computing losses."
            raise RuntimeError(msg)
        style_losses, content_losses = [], []
        for j, blk in enumerate(self.blocks):
            x = blk(x)
            if j in self.style_ids:
                tgt = self.style_targets[self.style_ids.index(j)]
                style_losses.append(F.mse_loss(gram_matrix(x), tgt))       # core_model.py:262-264
            if j in self.content_ids:
                tgt = self.content_targets[self.content_ids.index(j)]
                content_losses.append(F.mse_loss(x, tgt))                  # core_model.py:294-295
        return style_losses, content_losses

    def taps(self, x: torch.Tensor) -> list[torch.Tensor]:
        """Outputs of every block (the tensors the losses read), for per-layer checks."""
        outs = []
        for blk in self.blocks:
            x = blk(x)
            outs.append(x)
        return outs


def initialize_input(content_img: torch.Tensor, method: str) -> torch.Tensor:
    """core_model.py:66-100."""
    if method == "content":
        out = content_img.clone()
    elif method == "random":
        out = torch.randn_like(content_img)
    elif method == "white":
        out = torch.ones_like(content_img)
    else:
        msg = f"Unsupported initialization method: {method}"
        raise ValueError(msg)
    return out.requires_grad_(True)  # noqa: FBT003


def closure_step(model, x: torch.Tensor, style_w: float, content_w: float):  # noqa: ANN001, ANN201
    """optimization.py:286-327 minus bookkeeping: zero grad, forward, weighted sum, backward."""
    x.grad = None
    style_losses, content_losses = model(x)
    zero = torch.zeros((), dtype=x.dtype)
    style_score = torch.stack(style_losses).sum() if style_losses else zero
    content_score = torch.stack(content_losses).sum() if content_losses else zero
    loss = style_w * style_score + content_w * content_score
    loss.backward()
    return style_score.detach(), content_score.detach(), loss.detach(), style_losses, content_losses


def frame_u8(x: torch.Tensor, *, normalize: bool, rounding: bool = False) -> np.ndarray:
    """image_io.py:118-152 (denormalise, nan_to_num, clamp) + optimization.py:446-452
    (HWC, *255, truncating uint8 cast); ``rounding`` = main.py:203-214's final-frame variant."""
    img = x.detach()
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        img = img * std + mean
    img = torch.nan_to_num(img, nan=0.0, posinf=1.0, neginf=0.0).clamp(0, 1)
    arr = img.squeeze(0).permute(1, 2, 0).cpu().numpy() * 255
    if rounding:
        arr = arr.round()
    return arr.astype("uint8")


@dataclass
class RunResult:
    """What one optimisation run produces (the runner's observable outputs)."""

    style: list[float] = field(default_factory=list)      # per accepted step
    content: list[float] = field(default_factory=list)
    total: list[float] = field(default_factory=list)
    layer_style: list[float] = field(default_factory=list)    # step-1 per-layer style losses
    layer_content: list[float] = field(default_factory=list)  # step-1 per-layer content losses
    first_grad: torch.Tensor | None = None                # x.grad after the first closure
    final: torch.Tensor | None = None
    frames: list[np.ndarray] = field(default_factory=list)
    elapsed_s: float = 0.0


def run(model, x: torch.Tensor, optimizer: torch.optim.Optimizer, steps: int, *,  # noqa: ANN001
        style_w: float, content_w: float, save_every: int = 0,
        normalize: bool = True) -> RunResult:
    """optimization.py:162-202 + :274-348: ``optimizer.step(closure)`` per step; losses recorded
    once per ACCEPTED step (last closure evaluation); a frame every ``save_every`` steps."""
    import time

    res = RunResult()
    last: list = []

    def closure() -> torch.Tensor:
        s, c, t, sl, cl = closure_step(model, x, style_w, content_w)
        last[:] = [s, c, t]
        if res.first_grad is None:
            res.first_grad = x.grad.detach().clone()
            res.layer_style = [float(v.detach()) for v in sl]
            res.layer_content = [float(v.detach()) for v in cl]
        return t

    t0 = time.perf_counter()
    for step in range(1, steps + 1):
        optimizer.step(closure)
        s, c, t = last
        res.style.append(float(s))
        res.content.append(float(c))
        res.total.append(float(t))
        if save_every and step % save_every == 0:
            res.frames.append(frame_u8(x, normalize=normalize))
    res.elapsed_s = time.perf_counter() - t0
    res.final = x.detach().clone()
    return res


def make_optimizer(name: str, x: torch.Tensor, lr: float) -> torch.optim.Optimizer:
    """'adam' -> torch.optim.Adam([x], lr) (the injection seam optimization.py:104-105 and
    tests/test_optimization.py:178 use); 'lbfgs' -> the reference default (core_model.py:344-349)."""
    if name == "adam":
        return torch.optim.Adam([x], lr=lr)
    if name == "lbfgs":
        return torch.optim.LBFGS([x], lr=lr, max_iter=1, max_eval=1)
    msg = f"unknown optimizer {name}"
    raise ValueError(msg)
