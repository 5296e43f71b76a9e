"""Image helpers either side of the hot path: load + preprocess, de-normalisation and the
timelapse frame readback.

``load_image`` / ``validate_image_dimensions`` / ``apply_transforms`` / ``load_image_to_tensor``
keep the reference's names, errors and results (image_io.py:24-115).  On a CUDA device the image
crosses PCIe as BYTES (pinned uint8 staging, non-blocking copy) and ONE kernel does ToTensor +
Normalize there -- the same IEEE operations in torchvision's order, so the tensor is bit-identical
to the reference's fp32 upload at a quarter of the transfer.

``denormalize`` / ``prepare_image_for_output`` keep the reference's semantics (image_io.py:118-152).
``FrameReadback`` replaces the reference's synchronous ``.cpu().numpy() * 255 -> astype(uint8)``
(optimization.py:438-452) with ONE fused kernel (denormalise, nan_to_num, clamp, x255, truncate,
NCHW->HWC) writing 3 bytes per pixel, followed by a pinned-memory D2H copy on a side stream, so
the compute stream never waits for the copy.  The bytes are identical to the reference's.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from .constants import (DENORM_VIEW_SHAPE, IMAGENET_MEAN, IMAGENET_STD, MAX_DIMENSION,
                        MIN_DIMENSION)
from .logging_utils import logger


def load_image(path: str):  # noqa: ANN201
    """PIL image in RGB mode (reference image_io.py:24-46)."""
    from PIL import Image

    try:
        return Image.open(path).convert("RGB")
    except FileNotFoundError as exc:
        msg = f"Image file not found: '{path}'"
        raise FileNotFoundError(msg) from exc
    except OSError as exc:
        msg = f"Error loading image '{path}': {exc!s}"
        raise OSError(msg) from exc


def validate_image_dimensions(img) -> None:  # noqa: ANN001
    """Reject images below the minimum size, warn above the maximum (reference image_io.py:49-62)."""
    if img.width < MIN_DIMENSION or img.height < MIN_DIMENSION:
        msg = (f"Image too small: {img.width}x{img.height}. "
               f"Minimum dimension is {MIN_DIMENSION}px.")
        raise ValueError(msg)
    if img.width > MAX_DIMENSION or img.height > MAX_DIMENSION:
        logger.warning("Image is large: %dx%d. This may slow processing.", img.width, img.height)


# One page-locked staging buffer per image shape, reused by every load (pinning a fresh buffer per
# image costs milliseconds of cudaHostAlloc each time; configs[3] loads 128 images per run).
_PINNED_STAGING: dict[tuple[int, int], torch.Tensor] = {}


def _pinned_staging(height: int, width: int) -> torch.Tensor:
    buf = _PINNED_STAGING.get((height, width))
    if buf is None:
        if len(_PINNED_STAGING) >= 4:   # bounded: drop the oldest shape
            _PINNED_STAGING.pop(next(iter(_PINNED_STAGING)))
        buf = torch.empty(height, width, 3, dtype=torch.uint8).pin_memory()
        _PINNED_STAGING[(height, width)] = buf
    return buf


def apply_transforms(img, device: torch.device, *, normalize: bool) -> torch.Tensor:  # noqa: ANN001
    """``ToTensor`` (+ ``Normalize``) of a PIL image -> ``[1, 3, H, W]`` fp32 on ``device``
    (reference image_io.py:64-84)."""
    arr = torch.from_numpy(np.asarray(img.convert("RGB"), dtype=np.uint8).copy())   # [H, W, 3] u8
    if device.type == "cuda":
        staging = _pinned_staging(int(arr.shape[0]), int(arr.shape[1]))
        staging.copy_(arr)
        staged = staging.to(device, non_blocking=True)
        out = ops.image_from_u8(staged, normalize=normalize)
        # the staging buffer is reused by the next load: the copy must have left it
        torch.cuda.current_stream(device).synchronize()
        return out
    chw = arr.permute(2, 0, 1).contiguous().to(torch.float32).div(255)             # ToTensor
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
        chw = chw.sub_(mean).div_(std)                                             # Normalize
    return chw.unsqueeze(0).to(device)


def load_image_to_tensor(path: str, device: torch.device, *, normalize: bool = False) -> torch.Tensor:
    """Load, validate and preprocess an image as-is, no resizing (reference image_io.py:87-115)."""
    img = load_image(path)
    validate_image_dimensions(img)
    return apply_transforms(img, device, normalize=normalize)


def denormalize(tensor: torch.Tensor) -> torch.Tensor:
    """Undo the ImageNet normalisation (reference image_io.py:118-126)."""
    mean = torch.tensor(IMAGENET_MEAN).view(*DENORM_VIEW_SHAPE).to(tensor.device)
    std = torch.tensor(IMAGENET_STD).view(*DENORM_VIEW_SHAPE).to(tensor.device)
    return tensor * std + mean


def prepare_image_for_output(tensor: torch.Tensor, *, normalize: bool) -> torch.Tensor:
    """Float image in [0, 1] ready for saving (reference image_io.py:129-152)."""
    img = denormalize(tensor) if normalize else tensor
    img = torch.nan_to_num(img, nan=0.0, posinf=1.0, neginf=0.0)
    return img.clamp(0, 1)


@dataclass
class _Ticket:
    slot: int
    done: torch.cuda.Event
    tag: object


class FrameReadback:
    """Ring of (device u8 frame, pinned host frame) pairs + a copy stream."""

    def __init__(self, device: torch.device, height: int, width: int, depth: int = 3) -> None:
        self.device = device
        self.shape = (height, width, 3)
        self.depth = depth
        self._dev = [torch.empty(self.shape, device=device, dtype=torch.uint8)
                     for _ in range(depth)]
        self._host = [torch.empty(self.shape, dtype=torch.uint8).pin_memory()
                      for _ in range(depth)]
        self._stream = torch.cuda.Stream(device=device)
        self._next = 0
        self._inflight: deque[_Ticket] = deque()
        self.bytes_per_frame = height * width * 3

    @property
    def pending(self) -> int:
        return len(self._inflight)

    def submit(self, img: torch.Tensor, *, normalize: bool, tag: object = None,
               rounding: bool = False) -> None:
        """Enqueue conversion (current stream) + D2H (copy stream) of ``img`` [1,3,H,W]."""
        if len(self._inflight) >= self.depth:
            msg = "FrameReadback ring is full: collect() before submitting more frames"
            raise RuntimeError(msg)
        slot = self._next
        self._next = (slot + 1) % self.depth
        src = img.detach()
        ops.frame_to_u8(src if src.is_contiguous() else src.contiguous(), self._dev[slot],
                        denormalize=normalize, rounding=rounding)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        done = torch.cuda.Event()
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            self._host[slot].copy_(self._dev[slot], non_blocking=True)
            done.record(self._stream)
        self._inflight.append(_Ticket(slot=slot, done=done, tag=tag))

    def collect(self) -> tuple[np.ndarray, object]:
        """Oldest submitted frame as an owned ``[H, W, 3]`` uint8 array (+ its tag)."""
        ticket = self._inflight.popleft()
        ticket.done.synchronize()
        return self._host[ticket.slot].numpy().copy(), ticket.tag


_RING_CACHE: dict[tuple[int, int, int], FrameReadback] = {}


def shared_readback(device: torch.device, height: int, width: int) -> FrameReadback:
    """Process-wide readback ring per (device, H, W): pinned allocations cost milliseconds to
    hundreds of milliseconds, so consecutive jobs of the same size reuse one ring."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, height, width)
    ring = _RING_CACHE.get(key)
    if ring is None or ring.pending:
        ring = FrameReadback(device, height, width)
        _RING_CACHE[key] = ring
    return ring


def frame_to_numpy(img: torch.Tensor, *, normalize: bool, rounding: bool = False) -> np.ndarray:
    """One-shot synchronous variant (fused kernel + blocking copy)."""
    h, w = int(img.shape[-2]), int(img.shape[-1])
    out = torch.empty(h, w, 3, device=img.device, dtype=torch.uint8)
    src = img.detach()
    ops.frame_to_u8(src if src.is_contiguous() else src.contiguous(), out, denormalize=normalize,
                    rounding=rounding)
    return out.cpu().numpy()
