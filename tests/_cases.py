"""Golden-case plumbing shared by the CPU oracle tests, the GPU parity tests and bench tools.

A golden fixture (tests/golden/<name>.npz, written by oracle/make_golden.py from the real
reference) stores only outputs; inputs are regenerated here from the seeds in its config.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from oracle import stv_oracle as orc

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


def golden_names() -> list[str]:
    return sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))


def load_golden(name: str) -> tuple[dict, dict]:
    data = np.load(GOLDEN_DIR / f"{name}.npz", allow_pickle=False)
    cfg = json.loads(str(data["config"]))
    return cfg, {k: data[k] for k in data.files if k != "config"}


def case_inputs(cfg: dict) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor | None]:
    """(content, style, explicit initial image or None) -- must mirror oracle/make_golden.py."""
    norm = cfg.get("normalize", True)
    scale = cfg.get("scale", 1.0)
    content = orc.synthetic_image(1, cfg["h"], cfg["w"], normalize=norm, scale=scale)
    style = orc.synthetic_image(2, cfg.get("sh", cfg["h"]), cfg.get("sw", cfg["w"]),
                                normalize=norm, scale=scale)
    init = None
    if cfg["init"] == "random":
        gen = torch.Generator().manual_seed(3)
        init = torch.randn(content.shape, generator=gen) * scale
    elif cfg.get("init_noise"):  # content image + seeded Gaussian noise (explicit start)
        gen = torch.Generator().manual_seed(3)
        init = content + cfg["init_noise"] * scale * torch.randn(content.shape, generator=gen)
    return content, style, init


def initial_image(cfg: dict, content: torch.Tensor, init: torch.Tensor | None) -> torch.Tensor:
    if init is not None:
        return init.clone()
    if cfg["init"] == "content":
        return content.clone()
    if cfg["init"] == "white":
        return torch.ones_like(content)
    raise ValueError(cfg["init"])


def sample_stride(cfg: dict) -> int:
    """Spatial stride of the image-shaped samples a fixture stores (1 = everything); must mirror
    oracle/make_golden.py."""
    px = cfg["h"] * cfg["w"]
    if px > 1024 * 1024:
        return 8
    return 4 if px > 128 * 128 else 1


def subsample_like_golden(cfg: dict, t: torch.Tensor | np.ndarray) -> np.ndarray:
    """Big cases store strided samples [..., ::k, ::k] of image-shaped tensors."""
    arr = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t
    k = sample_stride(cfg)
    return np.ascontiguousarray(arr[..., ::k, ::k]) if k > 1 else arr


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def cosine(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
