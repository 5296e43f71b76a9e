"""Drop-in switch for the reference's orchestration.

The reference reaches its hot path through two module attributes of ``main.py``:
``stv_core_model.prepare_model_and_input`` (main.py:72-77) and ``stv_optimizer.OptimizationRunner``
(main.py:122-132) -- the seam its own tests use to swap them (tests/test_main.py:110-114,
977-1004).  ``enable`` points both at this package; everything else (config / TOML / CLI, image
loading, video writers, intro / outro, plots) keeps running from the reference.
"""
from __future__ import annotations

from types import ModuleType

from . import core_model as _core_model
from . import optimization as _optimization

_saved: dict[int, tuple[object, object]] = {}


def enable(stv_main: ModuleType) -> None:
    """Route ``stv_main.style_transfer`` (the reference's ``style_transfer_visualizer.main``)
    through the sm_100a kernels.  CUDA sm_100 only: a CPU run raises, it does not fall back."""
    key = id(stv_main)
    if key not in _saved:
        _saved[key] = (stv_main.stv_core_model.prepare_model_and_input,
                       stv_main.stv_optimizer.OptimizationRunner)
    stv_main.stv_core_model.prepare_model_and_input = _core_model.prepare_model_and_input
    stv_main.stv_optimizer.OptimizationRunner = _optimization.OptimizationRunner


def disable(stv_main: ModuleType) -> None:
    """Undo ``enable``."""
    saved = _saved.pop(id(stv_main), None)
    if saved is not None:
        stv_main.stv_core_model.prepare_model_and_input, \
            stv_main.stv_optimizer.OptimizationRunner = saved
