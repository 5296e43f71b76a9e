"""Device-side loss ring buffers with host syncs only at the logging cadence.

Behavioural mirror of the reference's LossAccumulator (loss_accumulator.py:26-213): per-step
losses stay on the device in fixed-capacity circular buffers; Python floats are materialised only
every ``log_every`` steps (or when forced); ``export_history`` returns the retained window in
chronological order.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

DEFAULT_HISTORY_CAPACITY = 2048


@dataclass(slots=True)
class LoggedLoss:
    """Loss scalars that have been synced to the host."""

    step: int
    style_loss: float
    content_loss: float
    total_loss: float


_NAMES = ("style_loss", "content_loss", "total_loss")


class LossAccumulator:
    """Aggregate loss tensors on the device; batch the device->host reads."""

    def __init__(self, *, log_every: int, history_capacity: int | None, track_history: bool,
                 device: torch.device, dtype: torch.dtype) -> None:
        self._log_every = max(1, log_every)
        self._history_capacity = max(1, history_capacity or DEFAULT_HISTORY_CAPACITY)
        self._track_history = track_history
        self._device = device
        self._buffer_dtype = torch.float16 if dtype == torch.float16 else torch.float32
        # one [3, capacity] ring: rows = style, content, total
        self._ring: torch.Tensor | None = None
        if track_history:
            self._ring = torch.empty(3, self._history_capacity, dtype=self._buffer_dtype,
                                     device=device)
        self._cursor = 0        # next slot to write
        self._filled = 0        # valid entries (<= capacity)
        self._records = 0       # total writes ever
        self._pending: tuple[int, torch.Tensor, torch.Tensor, torch.Tensor] | None = None
        self._last_logged: LoggedLoss | None = None
        # rows [capacity, 3] written by the captured step graph itself (see adopt_device_rows)
        self._device_rows: torch.Tensor | None = None

    def adopt_device_rows(self, rows: torch.Tensor) -> None:
        """Use ``rows`` ``[capacity, 3]`` as the history ring: the producer (the captured step graph,
        ``stv_step_scores``) writes row ``k % capacity`` for its k-th step by itself, so
        ``accumulate`` only does host bookkeeping -- no per-step device work at all.  Must be called
        before the first ``accumulate``; the producer's row counter must start at zero with it."""
        if self._records:
            msg = "adopt_device_rows must precede the first accumulate()"
            raise RuntimeError(msg)
        if rows.dim() != 2 or rows.shape[1] != 3 or rows.shape[0] != self._history_capacity:
            msg = (f"device rows must have shape [{self._history_capacity}, 3], "
                   f"got {tuple(rows.shape)}")
            raise ValueError(msg)
        self._device_rows = rows
        self._ring = None

    @property
    def capacity(self) -> int:
        return self._history_capacity

    @property
    def tracks_history(self) -> bool:
        return self._track_history

    @property
    def history_truncated(self) -> bool:
        """True once older entries have been overwritten."""
        return self._records > self._history_capacity

    def accumulate(self, step_idx: int, style_loss: torch.Tensor, content_loss: torch.Tensor,
                   total_loss: torch.Tensor, *, force: bool = False) -> LoggedLoss | None:
        """Record one step; return host scalars only on ``log_every`` steps (or ``force``)."""
        vals = (style_loss.detach(), content_loss.detach(), total_loss.detach())
        self._pending = (step_idx, *vals)
        if self._device_rows is not None:
            self._advance()  # the row is already on the device
        elif self._track_history:
            self._push(vals)
        if force or step_idx % self._log_every == 0:
            return self._sync_pending()
        return None

    def latest(self) -> LoggedLoss | None:
        return self._last_logged

    def export_history(self) -> dict[str, list[float]]:
        """Retained history, oldest first, as plain lists."""
        ring = self._device_rows.t() if self._device_rows is not None else self._ring
        if not self._track_history or self._filled == 0 or ring is None:
            return {name: [] for name in _NAMES}
        start = (self._cursor - self._filled) % self._history_capacity
        order = (torch.arange(self._filled, device=ring.device) + start) \
            % self._history_capacity
        window = ring.index_select(1, order).cpu()
        return {name: window[row].tolist() for row, name in enumerate(_NAMES)}

    def _push(self, vals: tuple[torch.Tensor, torch.Tensor, torch.Tensor]) -> None:
        if self._ring is None:
            msg = "History buffers are uninitialized."
            raise RuntimeError(msg)
        slot = self._cursor
        for row, val in enumerate(vals):
            self._ring[row, slot] = val.to(dtype=self._buffer_dtype, device=self._device)
        self._advance()

    def _advance(self) -> None:
        self._cursor = (self._cursor + 1) % self._history_capacity
        self._filled = min(self._filled + 1, self._history_capacity)
        self._records += 1

    def _sync_pending(self) -> LoggedLoss | None:
        if self._pending is None:
            return None
        step, style, content, total = self._pending
        if self._device_rows is not None:
            # the three scalars of the row just written, in ONE device->host copy
            row = self._device_rows[(self._cursor - 1) % self._history_capacity].tolist()
            logged = LoggedLoss(step=step, style_loss=float(row[0]), content_loss=float(row[1]),
                                total_loss=float(row[2]))
        else:
            logged = LoggedLoss(step=step, style_loss=self._to_float(style),
                                content_loss=self._to_float(content),
                                total_loss=self._to_float(total))
        self._last_logged = logged
        return logged

    def _to_float(self, tensor: torch.Tensor) -> float:
        return float(tensor.item())
