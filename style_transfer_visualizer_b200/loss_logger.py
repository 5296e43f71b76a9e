"""CSV loss log: ``step,style_loss,content_loss,total_loss`` rows at a fixed cadence.

Same file format and flushing behaviour as the reference's LossCSVLogger (loss_logger.py:14-126):
header written and flushed on open, one flushed row per logged step, idempotent ``close``.
"""
from __future__ import annotations

import csv
from pathlib import Path
from types import TracebackType

HEADER = ("step", "style_loss", "content_loss", "total_loss")


class LossCSVLogger:
    """Append loss rows to a CSV file every ``log_every`` steps."""

    def __init__(self, path: str | Path, log_every: int) -> None:
        self.path = Path(path)
        self.log_every = log_every
        self.path.parent.mkdir(parents=True, exist_ok=True)
        self.file = self.path.open("w", newline="", encoding="utf-8")  # OSError propagates
        self.writer = csv.writer(self.file)
        self.writer.writerow(list(HEADER))
        self.file.flush()

    def log(self, step: int, style_loss: float, content_loss: float, total_loss: float) -> None:
        """Write one row when ``step`` falls on the logging cadence."""
        if not self.writer or step % self.log_every != 0:
            return
        self.writer.writerow([step, style_loss, content_loss, total_loss])
        self.file.flush()

    def close(self) -> None:
        if self.file and not self.file.closed:
            self.file.close()

    def __enter__(self) -> "LossCSVLogger":
        return self

    def __exit__(self, exc_type: type[BaseException] | None, exc: BaseException | None,
                 tb: TracebackType | None) -> None:
        self.close()
