"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  For every case below the reference's own ``prepare_model_and_input``
(core_model.py:331-350) and ``OptimizationRunner.run`` (optimization.py:162-202) are executed on
the CPU with ``initialize_vgg`` patched to the seeded random-init VGG19 (pretrained weights are
unavailable offline) and an in-memory frame sink, and the observable outputs are stored:
per-layer losses and ``input_img.grad`` of the first closure, the per-step loss history returned
by ``run()``, the final image, the timelapse frames and the CSV text.

Inputs are NOT stored: they are regenerated from seeds by ``oracle.stv_oracle.synthetic_image`` /
``vgg19_features`` (same code on every box), which keeps the fixtures small.

    python -m oracle.make_golden            # writes tests/golden/<case>.npz
"""
from __future__ import annotations

import io
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_shim  # noqa: E402
from oracle.stv_oracle import synthetic_image  # noqa: E402

GOLDEN_DIR = ROOT / "tests" / "golden"
WEIGHT_SEED = 0

# name -> configuration.  `scale` multiplies the (normalised) images: 30x trips the 5e5 clamp.
CASES: dict[str, dict] = {
    "adam_content_64": dict(h=64, w=64, opt="adam", lr=0.01, steps=6, init="content",
                            save_every=2),
    "adam_random_64": dict(h=64, w=64, opt="adam", lr=0.05, steps=4, init="random"),
    "adam_white_odd_70x94": dict(h=70, w=94, opt="adam", lr=0.01, steps=3, init="white"),
    "adam_clamp_64": dict(h=64, w=64, opt="adam", lr=0.01, steps=3, init="random", scale=30.0),
    "adam_layers_024_13_64": dict(h=64, w=64, opt="adam", lr=0.01, steps=3, init="random",
                                  style_layers=[0, 2, 4], content_layers=[1, 3]),
    "adam_stylesize_96x128": dict(h=96, w=128, sh=80, sw=112, opt="adam", lr=0.01, steps=3,
                                  init="content"),
    "adam_nonorm_64": dict(h=64, w=64, opt="adam", lr=0.01, steps=3, init="content",
                           normalize=False, save_every=1),
    "lbfgs_random_64": dict(h=64, w=64, opt="lbfgs", lr=1.0, steps=8, init="random",
                            style_w=1e9),
    "lbfgs_content_64": dict(h=64, w=64, opt="lbfgs", lr=1.0, steps=6, init="content",
                             style_w=1e10),
    # L-BFGS (the reference's default optimiser) on a well-conditioned start: content image +
    # noise, style and content terms of similar size.  Chosen with a TF32-emulating oracle so that
    # TF32-level gradient differences do not fork the trajectory within 8 steps (the third step
    # -- the first one built from a single, tiny curvature pair -- is the sensitive one).
    "lbfgs_noisy_64": dict(h=64, w=64, opt="lbfgs", lr=1.0, steps=8, init="content",
                           init_noise=0.3, style_w=1e8, content_w=10.0),
    # BASELINE.json configs[1] at full size (first 8 of its 300 steps; the CPU needs ~1 s per step)
    "adam_content_512_c2": dict(h=512, w=512, opt="adam", lr=0.01, steps=8, init="content"),
    # BASELINE.json configs[0]: 256x256, Adam, 50 steps, content init, on CPU
    "adam_content_256_c1": dict(h=256, w=256, opt="adam", lr=0.01, steps=50, init="content",
                                save_every=25, csv=True),
    # BASELINE.json configs[2] at full size: 1920x1080 (pools floor 135 -> 67, the 5e5 Gram clamp
    # is live on conv1_1), first 2 of its 500 steps with a frame every step (~3 s per CPU step)
    "adam_content_1080p_c3": dict(h=1080, w=1920, opt="adam", lr=0.01, steps=2, init="content",
                                  save_every=1),
    # BASELINE.json configs[4] size: 3840x2160, first closure + one Adam step; random init so that
    # the content loss and its gradient are non-zero (the sharded path is checked against this)
    "adam_random_4k_c5": dict(h=2160, w=3840, opt="adam", lr=0.01, steps=1, init="random"),
}


class MemorySink:
    """In-memory VideoFrameSink (video.py:117-126 protocol), as the reference tests use."""

    def __init__(self) -> None:
        self.frames: list[np.ndarray] = []
        self._size = None

    def append_data(self, frame: np.ndarray) -> None:
        self.frames.append(np.array(frame, copy=True))

    def close(self) -> None:
        pass


def case_inputs(cfg: dict) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor | None]:
    """content, style, explicit initial image (None when init is deterministic)."""
    norm = cfg.get("normalize", True)
    scale = cfg.get("scale", 1.0)
    content = synthetic_image(1, cfg["h"], cfg["w"], normalize=norm, scale=scale)
    style = synthetic_image(2, cfg.get("sh", cfg["h"]), cfg.get("sw", cfg["w"]), normalize=norm,
                            scale=scale)
    init = None
    if cfg["init"] == "random":
        gen = torch.Generator().manual_seed(3)
        init = torch.randn(content.shape, generator=gen) * scale
    elif cfg.get("init_noise"):  # content image + seeded Gaussian noise (explicit start)
        gen = torch.Generator().manual_seed(3)
        init = content + cfg["init_noise"] * scale * torch.randn(content.shape, generator=gen)
    return content, style, init


def run_reference(name: str, cfg: dict) -> dict:
    ref = reference_shim.load()
    reference_shim.patch_random_vgg(ref, WEIGHT_SEED)
    content, style, init = case_inputs(cfg)
    style_layers = cfg.get("style_layers", [0, 5, 10, 19, 28])
    content_layers = cfg.get("content_layers", [21])
    norm = cfg.get("normalize", True)
    tmp = tempfile.mkdtemp()
    csv_path = str(Path(tmp) / "loss.csv") if cfg.get("csv") else None
    config = ref.config.StyleTransferConfig.model_validate({
        "optimization": {
            "steps": cfg["steps"], "style_w": cfg.get("style_w", 1e5),
            "content_w": cfg.get("content_w", 1.0), "lr": cfg["lr"],
            "init_method": cfg["init"], "normalize": norm, "style_layers": style_layers,
            "content_layers": content_layers,
        },
        "video": {"save_every": cfg.get("save_every") or cfg["steps"] + 1},
        "output": {"log_every": 1, "log_loss": csv_path},
    })
    model, input_img, lbfgs = ref.core_model.prepare_model_and_input(
        content, style, torch.device("cpu"), config.optimization)
    if init is not None:  # same explicit start for every implementation (RNG streams differ per device)
        with torch.no_grad():
            input_img.copy_(init)
    optimizer = lbfgs if cfg["opt"] == "lbfgs" else torch.optim.Adam([input_img], lr=cfg["lr"])

    # first-closure observables, computed on a detached copy with the reference's own model
    probe = input_img.detach().clone().requires_grad_(True)
    sl, cl = model(probe)
    total = config.optimization.style_w * torch.stack(sl).sum() \
        + config.optimization.content_w * torch.stack(cl).sum()
    total.backward()

    sink = MemorySink() if cfg.get("save_every") else None

    class _Bar:
        def update(self, n=1): ...
        def set_postfix(self, *a, **k): ...
        def close(self): ...

    runner = ref.optimization.OptimizationRunner(
        model, input_img, config, optimizer=optimizer, progress_bar=_Bar(), video_writer=sink)
    final, history, _elapsed = runner.run()

    if csv_path:  # CSV logging replaces the in-memory history (optimization.py:194-202)
        text = Path(csv_path).read_text(encoding="utf-8")
        rows = [r.split(",") for r in text.strip().splitlines()[1:]]
        history = {"style_loss": [float(r[1]) for r in rows],
                   "content_loss": [float(r[2]) for r in rows],
                   "total_loss": [float(r[3]) for r in rows]}
    else:
        text = ""
    out = {
        "config": json.dumps({**cfg, "style_layers": style_layers,
                              "content_layers": content_layers, "weight_seed": WEIGHT_SEED,
                              "torch": torch.__version__}),
        "layer_style": np.array([float(v) for v in sl], dtype=np.float64),
        "layer_content": np.array([float(v) for v in cl], dtype=np.float64),
        "first_grad": probe.grad.numpy().astype(np.float32),
        "style_loss": np.array(history["style_loss"], dtype=np.float64),
        "content_loss": np.array(history["content_loss"], dtype=np.float64),
        "total_loss": np.array(history["total_loss"], dtype=np.float64),
        "final": final.detach().numpy().astype(np.float32),
        "csv": np.array(text),
        "style_targets_sum": np.array([float(t.double().sum()) for t in model.style_targets]),
        "style_targets_absmax": np.array([float(t.abs().max()) for t in model.style_targets]),
    }
    if sink is not None:
        out["frames"] = np.stack(sink.frames) if sink.frames else np.zeros((0,), dtype=np.uint8)
    px = cfg["h"] * cfg["w"]
    k = 8 if px > 1024 * 1024 else (4 if px > 128 * 128 else 1)  # tests/_cases.py::sample_stride
    if k > 1:  # keep big cases small: strided samples + checksums of the full tensors
        for key in ("first_grad", "final"):
            full = out[key]
            out[key + "_sum"] = np.array(float(full.astype(np.float64).sum()))
            out[key + "_l2"] = np.array(float(np.sqrt((full.astype(np.float64) ** 2).sum())))
            out[key] = np.ascontiguousarray(full[..., ::k, ::k])
        if "frames" in out and out["frames"].size:
            out["frames"] = np.ascontiguousarray(out["frames"][:, ::k, ::k, :])
    return out


def main() -> None:
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(1234)
        out = run_reference(name, cfg)
        buf = io.BytesIO()
        np.savez_compressed(buf, **out)
        (GOLDEN_DIR / f"{name}.npz").write_bytes(buf.getvalue())
        print(f"{name}: total_loss[0]={out['total_loss'][0]:.6e} -> [-1]={out['total_loss'][-1]:.6e} "
              f"|grad|max={np.abs(out['first_grad']).max():.3e} ({len(buf.getvalue()) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
