"""Shared ``style_transfer`` logger (same logger name as the reference's logging_utils.py:51 so
that log capture configured for the reference keeps working)."""
from __future__ import annotations

import logging


def setup_logger(name: str = "style_transfer", level: int = logging.INFO) -> logging.Logger:
    log = logging.getLogger(name)
    log.setLevel(level)
    if not log.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter("%(asctime)s [%(levelname)s] %(message)s"))
        log.addHandler(handler)
        log.propagate = False
    return log


logger = setup_logger()
