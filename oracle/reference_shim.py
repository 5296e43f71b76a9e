"""Import the UNMODIFIED reference from /root/reference/src (build container only).

TEST INFRASTRUCTURE.  The reference needs ``tomlkit`` and ``imageio`` at import time
(config.py:12, video.py:13); neither is installed here and neither touches the hot path, so two
in-memory stubs stand in for them.  ``initialize_vgg`` (core_model.py:103-117) would download
pretrained weights; it is patched to the seeded random-init network exactly as the reference's
own tests patch it (tests/test_core_model.py:149-157).  Nothing here is importable on the GPU box
(no /root/reference there): callers must check ``available()`` first.
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path


def reference_src() -> Path | None:
    for cand in (os.environ.get("STV_REFERENCE_SRC"), "/root/reference/src"):
        if cand and (Path(cand) / "style_transfer_visualizer" / "core_model.py").exists():
            return Path(cand)
    return None


def available() -> bool:
    return reference_src() is not None


def load():  # noqa: ANN201
    """Return the reference modules (core_model, optimization, config, loss_accumulator,
    loss_logger, image_io) imported from the read-only reference tree."""
    src = reference_src()
    if src is None:
        msg = "reference source tree not found (set STV_REFERENCE_SRC)"
        raise RuntimeError(msg)
    if "tomlkit" not in sys.modules:
        import tomllib

        stub = types.ModuleType("tomlkit")
        stub.load = lambda f: tomllib.loads(f.read())
        sys.modules["tomlkit"] = stub
    if "imageio" not in sys.modules:
        iio = types.ModuleType("imageio")
        plugins = types.ModuleType("imageio.plugins")
        ffmpeg = types.ModuleType("imageio.plugins.ffmpeg")
        iio.plugins = plugins
        plugins.ffmpeg = ffmpeg
        sys.modules["imageio"] = iio
        sys.modules["imageio.plugins"] = plugins
        sys.modules["imageio.plugins.ffmpeg"] = ffmpeg
    if str(src) not in sys.path:
        sys.path.insert(0, str(src))
    import style_transfer_visualizer.config as config
    import style_transfer_visualizer.core_model as core_model
    import style_transfer_visualizer.image_io as image_io
    import style_transfer_visualizer.loss_accumulator as loss_accumulator
    import style_transfer_visualizer.loss_logger as loss_logger
    import style_transfer_visualizer.optimization as optimization

    return types.SimpleNamespace(core_model=core_model, optimization=optimization, config=config,
                                 loss_accumulator=loss_accumulator, loss_logger=loss_logger,
                                 image_io=image_io)


def patch_random_vgg(ref, seed: int) -> None:  # noqa: ANN001
    """Replace the pretrained-weight loader by the seeded random-init network."""
    from oracle.stv_oracle import vgg19_features

    ref.core_model.initialize_vgg = lambda: vgg19_features(seed)
