"""Synthetic inputs for benchmarks and demos: seeded random-init VGG19 features (pretrained
weights need a download) and uniform-random images passed through the load transform
(ToTensor + ImageNet Normalize, reference image_io.py:64-84)."""
from __future__ import annotations

import torch
from torch import nn

from .constants import IMAGENET_MEAN, IMAGENET_STD


def random_vgg19_features(seed: int) -> nn.Module:
    """``torchvision.models.vgg19(weights=None).features`` under ``torch.manual_seed(seed)``,
    eval mode, frozen -- what the reference's tests substitute for ``initialize_vgg``."""
    from torchvision.models import vgg19

    torch.manual_seed(seed)
    feats = vgg19(weights=None).features.eval()
    for p in feats.parameters():
        p.requires_grad_(False)  # noqa: FBT003
    return feats


def synthetic_image(seed: int, height: int, width: int, *, normalize: bool = True) -> torch.Tensor:
    """``[1, 3, H, W]`` CPU tensor of uniform-random pixels, optionally ImageNet-normalised."""
    gen = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, height, width, generator=gen)
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
        img = (img - mean) / std
    return img
