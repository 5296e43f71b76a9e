"""The unmodified reference orchestration drives this package through its own seam.

``/root/reference/src/style_transfer_visualizer/main.py`` calls
``stv_core_model.prepare_model_and_input`` (main.py:72-77) and ``stv_optimizer.OptimizationRunner``
(main.py:122-132) through module attributes; ``integration.enable`` swaps both (INTEGRATION.md
section 1), exactly as the reference's tests swap them (tests/test_main.py:110-114, 977-1004).
Runs only where the reference tree is mounted (the build container); no GPU needed -- on a CPU the
call must arrive in this package and fail loudly there (no fallback).
"""
from __future__ import annotations

import inspect

import numpy as np
import pytest
import torch

from oracle import reference_shim

pytestmark = pytest.mark.skipif(not reference_shim.available(),
                                reason="reference tree not mounted")


@pytest.fixture
def ref_main():
    reference_shim.load()
    import style_transfer_visualizer.main as stv_main

    from style_transfer_visualizer_b200 import integration

    integration.enable(stv_main)
    try:
        yield stv_main
    finally:
        integration.disable(stv_main)


def test_enable_swaps_both_seam_attributes_and_restores_them() -> None:
    reference_shim.load()
    import style_transfer_visualizer.core_model as ref_model
    import style_transfer_visualizer.main as stv_main
    import style_transfer_visualizer.optimization as ref_opt

    import style_transfer_visualizer_b200.core_model as b200_model
    import style_transfer_visualizer_b200.optimization as b200_opt
    from style_transfer_visualizer_b200 import integration

    before = (ref_model.prepare_model_and_input, ref_opt.OptimizationRunner)
    integration.enable(stv_main)
    try:
        assert stv_main.stv_core_model.prepare_model_and_input is b200_model.prepare_model_and_input
        assert stv_main.stv_optimizer.OptimizationRunner is b200_opt.OptimizationRunner
    finally:
        integration.disable(stv_main)
    assert (stv_main.stv_core_model.prepare_model_and_input,
            stv_main.stv_optimizer.OptimizationRunner) == before


def test_call_sites_are_signature_compatible(ref_main) -> None:  # noqa: ANN001
    """Every argument main.py passes (main.py:72-77 positionally, main.py:122-132 by keyword) binds
    to this package's callables; nothing the reference passes is dropped or renamed."""
    reference_shim.load()
    import style_transfer_visualizer.core_model as ref_model  # patched attrs live on the modules
    import style_transfer_visualizer.optimization as ref_opt

    from style_transfer_visualizer_b200 import core_model as b200_model
    from style_transfer_visualizer_b200 import optimization as b200_opt

    # originals were saved by enable(); compare against the reference classes by name
    ref_prepare = inspect.signature(
        inspect.unwrap(getattr(ref_model, "prepare_model_and_input")))  # noqa: B009
    assert list(ref_prepare.parameters)[:4] == list(
        inspect.signature(b200_model.prepare_model_and_input).parameters)[:4] == \
        ["content_img", "style_img", "device", "optimization"]
    mine = inspect.signature(b200_opt.OptimizationRunner.__init__).parameters
    for name in ("model", "input_img", "config", "optimizer", "optimizer_factory", "progress_bar",
                 "callbacks", "video_writer", "gif_collector", "intro_last_frame",
                 "intro_crossfade_frames"):
        assert name in mine, name
    # the extra knobs of this package are keyword-only with defaults: reference call sites bind
    inspect.signature(b200_opt.OptimizationRunner.__init__).bind(
        None, "model", "input_img", "config", optimizer="opt", video_writer=None,
        gif_collector=None, intro_last_frame=None, intro_crossfade_frames=0)
    assert ref_opt.OptimizationRunner is b200_opt.OptimizationRunner  # seam is patched right now
    for attr in ("StepMetrics", "StepTensors", "OptimizationCallbacks", "ProgressReporter"):
        assert hasattr(b200_opt, attr)


def test_reference_style_transfer_reaches_this_package(ref_main, tmp_path) -> None:  # noqa: ANN001
    """The reference's own ``style_transfer()`` (validation, seeding, device setup, image loading,
    video-mode selection -- all reference code) hands over to this package at main.py:72.  Without
    an sm_100 GPU the hand-over must end in this package's loud no-fallback error, which proves the
    call went through the seam and not into the reference's PyTorch model."""
    from PIL import Image

    from style_transfer_visualizer.config import StyleTransferConfig
    from style_transfer_visualizer.type_defs import InputPaths

    from style_transfer_visualizer_b200._native import NativeLibraryError

    if torch.cuda.is_available():
        pytest.skip("CPU-side wiring check")
    rng = np.random.default_rng(0)
    for name in ("content.png", "style.png"):
        Image.fromarray(rng.integers(0, 255, (64, 80, 3), dtype=np.uint8)).save(tmp_path / name)
    config = StyleTransferConfig.model_validate({
        "optimization": {"steps": 1},
        "video": {"create_video": False, "final_only": True},
        "hardware": {"device": "cpu"},
        "output": {"output": str(tmp_path / "out"), "plot_losses": False},
    })
    paths = InputPaths(content_path=str(tmp_path / "content.png"),
                       style_path=str(tmp_path / "style.png"))
    import style_transfer_visualizer_b200.core_model as b200_model

    original = b200_model.initialize_vgg
    from style_transfer_visualizer_b200.synthetic import random_vgg19_features

    b200_model.initialize_vgg = lambda: random_vgg19_features(0)  # no weight download offline
    try:
        with pytest.raises(NativeLibraryError, match="no CPU fallback"):
            ref_main.style_transfer(paths, config)
    finally:
        b200_model.initialize_vgg = original
