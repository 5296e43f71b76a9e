"""The slice of the reference's video module the optimisation loop itself touches: the frame-sink
protocol (video.py:117-126) and the intro crossfade (video.py:138-152, 260-274).  Encoding,
intro/outro synthesis and GIF assembly are CPU-side work outside this package's scope."""
from __future__ import annotations

from typing import Protocol

import numpy as np

from .constants import INTRO_MAX_CROSSFADE_FRAMES


class VideoFrameSink(Protocol):
    """Writer-like object: receives ``[H, W, 3]`` uint8 RGB frames."""

    _size: tuple[int, int] | None

    def append_data(self, frame: np.ndarray) -> None: ...

    def close(self) -> None: ...


def blend_frames(frame_a: np.ndarray, frame_b: np.ndarray, alpha: float) -> np.ndarray:
    """Rounded linear blend ``(1 - alpha) * a + alpha * b`` of two RGB frames."""
    if frame_a.shape != frame_b.shape:
        msg = "Frames must share shape for blending"
        raise ValueError(msg)
    mixed = frame_a.astype(np.float32) * (1.0 - alpha) + frame_b.astype(np.float32) * alpha
    return np.clip(np.rint(mixed), 0, 255).astype(np.uint8)


def append_crossfade(writer: VideoFrameSink, start_frame: np.ndarray, end_frame: np.ndarray,
                     frame_count: int, *, max_frames: int = INTRO_MAX_CROSSFADE_FRAMES) -> None:
    """Append ``min(frame_count, max_frames)`` blended frames between two frames."""
    if frame_count <= 0:
        return
    n = max(1, min(frame_count, max_frames))
    for k in range(n):
        writer.append_data(blend_frames(start_frame, end_frame, (k + 1) / (n + 1)))
