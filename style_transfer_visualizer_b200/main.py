"""Orchestration of one style-transfer job around the accelerated hot path.

Keeps the shape of the reference's ``main.style_transfer`` (main.py:20-167): seed, device, load and
normalise the two images, ``prepare_model_and_input``, ``OptimizationRunner.run()``, save the final
PNG, return the clamped image.  Video encoding, intro/outro synthesis, galleries and plots are the
reference's CPU-side subsystems and stay out of scope: frames are delivered to any
``VideoFrameSink`` the caller injects.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from pathlib import Path

import torch

from . import core_model as stv_core_model
from . import optimization as stv_optimizer
from .config import StyleTransferConfig
from .image_io import frame_to_numpy, load_image_to_tensor
from .logging_utils import logger
from .video import VideoFrameSink


@dataclass(slots=True)
class InputPaths:
    """Content and style input image paths (reference type_defs.py:24-29)."""

    content_path: str
    style_path: str


def setup_random_seed(seed: int) -> None:
    """Seed torch (CPU + CUDA) and ``random`` (reference runtime/device.py:31-42)."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    random.seed(seed)


def setup_device(device_name: str) -> torch.device:
    """Resolve the device.  Unlike the reference (runtime/device.py:12-28) there is no silent CPU
    fallback: the kernels exist for sm_100a only, so a missing GPU is an error."""
    if not device_name.startswith("cuda") or not torch.cuda.is_available():
        msg = (f"device '{device_name}' requested, but this implementation needs a CUDA sm_100a "
               "GPU (no CPU fallback)")
        raise RuntimeError(msg)
    device = torch.device(device_name)
    logger.info("Using device: %s", device)
    return device


def style_transfer(paths: InputPaths, config: StyleTransferConfig, *,
                   video_writer: VideoFrameSink | None = None) -> torch.Tensor:
    """Run one job end to end; returns ``input_img.detach().clamp(0, 1)`` like the reference."""
    for p in (paths.content_path, paths.style_path):
        if not Path(p).is_file():
            msg = f"Input image not found: {p}"
            raise FileNotFoundError(msg)
    if config.video.final_only:
        config.video.create_video = False
        config.video.create_gif = False
        config.video.save_every = config.optimization.steps + 1

    setup_random_seed(config.optimization.seed)
    device = setup_device(config.hardware.device)
    normalize = config.optimization.normalize
    content_img = load_image_to_tensor(paths.content_path, device, normalize=normalize)
    style_img = load_image_to_tensor(paths.style_path, device, normalize=normalize)

    model, input_img, optimizer = stv_core_model.prepare_model_and_input(
        content_img, style_img, device, config.optimization)

    out_dir = Path(config.output.output)
    out_dir.mkdir(parents=True, exist_ok=True)
    if config.video.create_video and video_writer is None:
        logger.warning("Video encoding is outside this package's scope; pass a VideoFrameSink "
                       "to receive timelapse frames. Continuing without a video.")

    runner = stv_optimizer.OptimizationRunner(model, input_img, config, optimizer=optimizer,
                                              video_writer=video_writer)
    input_img, _history, elapsed = runner.run()
    if video_writer is not None:
        video_writer.close()

    content_name, style_name = Path(paths.content_path).stem, Path(paths.style_path).stem
    final_path = out_dir / f"stylized_{content_name}_x_{style_name}.png"
    from PIL import Image

    Image.fromarray(frame_to_numpy(input_img, normalize=normalize, rounding=True)).save(final_path)
    logger.info("Style transfer completed in %.2f seconds", elapsed)
    logger.info("Final stylized image saved to: %s", final_path)
    return input_img.detach().clamp(0, 1)
