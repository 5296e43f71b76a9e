"""CPU tests of the host side: C-ABI symbol coverage, config / CLI, loss logging hooks and the
runner's behavioural contract (modelled on the reference's tests/test_optimization.py,
test_loss_accumulator.py, test_loss_logger.py, test_config.py).  No kernel is launched here."""
from __future__ import annotations

import csv
import logging
import re
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from style_transfer_visualizer_b200 import _native as nat
from style_transfer_visualizer_b200.config import (ConfigLoader, StyleTransferConfig,
                                                   build_config_from_cli, parse_int_list)
from style_transfer_visualizer_b200.loss_accumulator import LossAccumulator
from style_transfer_visualizer_b200.loss_logger import LossCSVLogger
from style_transfer_visualizer_b200.optimization import (OptimizationCallbacks,
                                                         OptimizationRunner, StepMetrics)

ROOT = Path(__file__).resolve().parent.parent


# --------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol() -> None:
    header = (ROOT / "include" / "stv_b200.h").read_text()
    declared = set(re.findall(r"\b(stv_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = nat.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in stv_b200.h but not exported"
    assert declared == set(nat.EXPORTED_SYMBOLS)
    assert lib.stv_abi_version() == 4
    assert lib.stv_reduce_scratch_floats() > 0
    assert lib.stv_gram_workspace_bytes(64 * 64, 64) > 0


def test_no_cpu_fallback() -> None:
    """The product path must fail loudly off-GPU (it never routes through the oracle)."""
    import style_transfer_visualizer_b200.core_model as cm
    from oracle import stv_oracle as orc

    with pytest.raises(nat.NativeLibraryError, match="no CPU fallback"):
        nat.require_device(torch.device("cpu"))
    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: orc.vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5], [2])
    finally:
        cm.initialize_vgg = original
    img = torch.zeros(1, 3, 64, 64)
    with pytest.raises(nat.NativeLibraryError):
        model.set_targets(img, img)
    with pytest.raises(RuntimeError, match="style_targets must be set"):
        model(img)
    src = "".join(p.read_text() for p in (ROOT / "style_transfer_visualizer_b200").glob("*.py"))
    assert "oracle" not in src.replace("oracle/", "")  # the package never imports the oracle


def test_feature_blocks_match_reference_layout() -> None:
    """Default taps cut vgg19.features into the 6 blocks of SURVEY section 3.3."""
    import style_transfer_visualizer_b200.core_model as cm
    from oracle import stv_oracle as orc

    blocks, content_ids, style_ids = cm.create_feature_blocks(
        orc.vgg19_features(0), [0, 5, 10, 19, 28], [21])
    assert [len(b) for b in blocks] == [1, 5, 5, 9, 2, 7]
    assert style_ids == [0, 1, 2, 3, 5] and content_ids == [4]
    assert not any(m.inplace for b in blocks for m in b if isinstance(m, nn.ReLU))
    x = torch.zeros(1, 3, 8, 8)
    with pytest.raises(ValueError, match="Unsupported initialization method"):
        cm.initialize_input(x, "bogus")  # type: ignore[arg-type]
    with pytest.raises(TypeError):
        cm.initialize_input("x", "content")  # type: ignore[arg-type]
    assert cm.initialize_input(x, "white").requires_grad
    assert torch.equal(cm.initialize_input(x + 2, "content").detach(), x + 2)


# --------------------------------------------------------------------------- config / CLI
def test_config_defaults_and_bounds() -> None:
    cfg = StyleTransferConfig.model_validate({})
    assert cfg.optimization.steps == 1500 and cfg.optimization.style_w == 1e5
    assert cfg.optimization.lr == 1.0 and cfg.optimization.init_method == "random"
    assert cfg.optimization.style_layers == [0, 5, 10, 19, 28]
    assert cfg.optimization.content_layers == [21]
    assert cfg.video.save_every == 20 and cfg.output.log_every == 10
    assert cfg.hardware.device == "cuda"
    with pytest.raises(ValueError):  # noqa: PT011
        StyleTransferConfig.model_validate({"optimization": {"steps": 0}})
    with pytest.raises(ValueError):  # noqa: PT011
        StyleTransferConfig.model_validate({"video": {"fps": 61}})


def test_cli_overrides_and_toml(tmp_path: Path, caplog: pytest.LogCaptureFixture) -> None:
    toml = tmp_path / "config.toml"
    toml.write_text("[optimization]\nsteps = 7\nstyle_w = 3.0\n[video]\nsave_every = 4\n"
                    "[output]\nlog_every = 2\n")
    base = ConfigLoader.load(str(toml))
    assert base.optimization.steps == 7 and base.video.save_every == 4
    cfg = build_config_from_cli({"config": str(toml), "steps": 9, "style_layers": "0,2,4",
                                 "no_normalize": True, "final_only": True})
    assert cfg.optimization.steps == 9 and cfg.optimization.style_w == 3.0
    assert cfg.optimization.style_layers == [0, 2, 4] and not cfg.optimization.normalize
    assert cfg.video.final_only and cfg.output.log_every == 2
    assert parse_int_list("1, 2,3") == [1, 2, 3] and parse_int_list([4]) == [4]
    with pytest.raises(FileNotFoundError):
        ConfigLoader.load(str(tmp_path / "missing.toml"))
    from style_transfer_visualizer_b200.logging_utils import logger

    logger.propagate = True
    try:
        with caplog.at_level(logging.WARNING, logger="style_transfer"):
            cfg = build_config_from_cli({"log_loss": str(tmp_path / "l.csv")})
    finally:
        logger.propagate = False
    assert not cfg.output.plot_losses
    assert "Loss plotting is disabled" in caplog.text
    from style_transfer_visualizer_b200.cli import build_arg_parser

    ns = build_arg_parser().parse_args(["--content", "c.png", "--style", "s.png", "--steps", "3"])
    assert vars(ns)["steps"] == 3 and "lr" not in vars(ns)


# --------------------------------------------------------------------------- logging hooks
def test_csv_logger_rows_and_cadence(tmp_path: Path) -> None:
    path = tmp_path / "sub" / "loss.csv"
    with LossCSVLogger(path, log_every=2) as lg:
        for step in range(1, 6):
            lg.log(step, 1.0, 0.5, 1.5)
    rows = list(csv.reader(path.open()))
    assert rows[0] == ["step", "style_loss", "content_loss", "total_loss"]
    assert rows[1:] == [["2", "1.0", "0.5", "1.5"], ["4", "1.0", "0.5", "1.5"]]
    lg.close()  # idempotent


def test_loss_accumulator_sync_cadence_and_ring() -> None:
    acc = LossAccumulator(log_every=3, history_capacity=4, track_history=True,
                          device=torch.device("cpu"), dtype=torch.float32)
    calls = []
    acc._to_float = lambda t: calls.append(1) or float(t)  # noqa: SLF001
    out = [acc.accumulate(i, torch.tensor(float(i)), torch.tensor(2.0 * i), torch.tensor(3.0 * i))
           for i in range(1, 8)]
    assert [o is not None for o in out] == [False, False, True, False, False, True, False]
    assert len(calls) == 6 and out[5].step == 6 and out[5].total_loss == 18.0
    hist = acc.export_history()
    assert hist["style_loss"] == [4.0, 5.0, 6.0, 7.0] and acc.history_truncated
    assert acc.accumulate(8, torch.tensor(1.0), torch.tensor(1.0), torch.tensor(1.0),
                          force=True).step == 8
    assert acc.latest().step == 8
    empty = LossAccumulator(log_every=1, history_capacity=None, track_history=False,
                            device=torch.device("cpu"), dtype=torch.float32)
    assert empty.export_history() == {"style_loss": [], "content_loss": [], "total_loss": []}
    assert empty.capacity == 2048 and not empty.tracks_history


# --------------------------------------------------------------------------- runner contract
class TinyModel(nn.Module):
    """Duck-typed model: forward(x) -> ([style], [content]) attached to x's graph."""

    def forward(self, x: torch.Tensor):  # noqa: ANN201
        return [(x ** 2).mean()], [(x - 1).abs().mean()]


class Bar:
    def __init__(self) -> None:
        self.updates = 0
        self.postfix = []
        self.closed = False

    def update(self, n=1):  # noqa: ANN001, ANN201
        self.updates += n

    def set_postfix(self, d=None, **k):  # noqa: ANN001, ANN003, ANN201
        self.postfix.append(d)

    def close(self):  # noqa: ANN201
        self.closed = True


class Sink:
    def __init__(self) -> None:
        self.frames = []
        self._size = None

    def append_data(self, f):  # noqa: ANN001, ANN201
        self.frames.append(f)

    def close(self):  # noqa: ANN201
        pass


class MultiProbeSGD(torch.optim.SGD):
    """Evaluates the closure several times per step (like a line-search optimiser)."""

    def step(self, closure=None):  # noqa: ANN001, ANN201
        for _ in range(3):
            loss = closure()
        super().step()
        return loss


def _cfg(steps: int = 6, **kw) -> StyleTransferConfig:  # noqa: ANN003
    return StyleTransferConfig.model_validate({
        "optimization": {"steps": steps, "normalize": kw.get("normalize", False)},
        "video": {"save_every": kw.get("save_every", 2)},
        "output": {"log_every": kw.get("log_every", 2), "log_loss": kw.get("log_loss")},
    })


def _img() -> torch.Tensor:
    return torch.rand(1, 3, 8, 8).requires_grad_(True)


def test_runner_history_frames_progress() -> None:
    x = _img()
    sink, bar = Sink(), Bar()
    seen: list[StepMetrics] = []
    frames_cb: list[int] = []
    starts: list[int] = []
    cbs = OptimizationCallbacks(on_step_start=starts.append, on_step_end=seen.append,
                                on_video_frame=lambda f, s: frames_cb.append(s))
    runner = OptimizationRunner(TinyModel(), x, _cfg(), optimizer=torch.optim.Adam([x], lr=0.1),
                                progress_bar=bar, callbacks=cbs, video_writer=sink)
    out, hist, elapsed = runner.run()
    assert out is x and elapsed >= 0
    assert all(len(hist[k]) == 6 for k in ("style_loss", "content_loss", "total_loss"))
    assert hist["total_loss"][-1] < hist["total_loss"][0]
    assert len(sink.frames) == 3 and frames_cb == [2, 4, 6] and starts == [1, 2, 3, 4, 5, 6]
    assert sink.frames[0].dtype == np.uint8 and sink.frames[0].shape == (8, 8, 3)
    assert bar.updates == 6 and not bar.closed  # a supplied bar is not owned
    assert [m.has_values for m in seen] == [False, True, False, True, False, True]
    # frame bytes follow the reference: clamp -> *255 -> truncation
    want = (x.detach().clamp(0, 1).squeeze(0).permute(1, 2, 0).numpy() * 255).astype("uint8")
    assert np.array_equal(sink.frames[-1], want)


def test_runner_multi_probe_one_frame_per_accepted_step() -> None:
    x = _img()
    sink, bar = Sink(), Bar()
    runner = OptimizationRunner(TinyModel(), x, _cfg(steps=4, save_every=1),
                                optimizer=MultiProbeSGD([x], lr=0.01), progress_bar=bar,
                                video_writer=sink)
    runner.run()
    assert runner._closure_calls == 12 and bar.updates == 4 and len(sink.frames) == 4  # noqa: SLF001


def test_runner_csv_logging_and_fallback(tmp_path: Path) -> None:
    x = _img()
    path = tmp_path / "loss.csv"
    runner = OptimizationRunner(TinyModel(), x, _cfg(log_loss=str(path)),
                                optimizer=torch.optim.Adam([x], lr=0.1), progress_bar=Bar())
    _, hist, _ = runner.run()
    assert hist == {}  # CSV replaces the in-memory history
    rows = list(csv.reader(path.open()))
    assert [r[0] for r in rows[1:]] == ["2", "4", "6"] and runner.loss_logger.file.closed
    errors: list[Exception] = []
    bad = tmp_path / "dir"
    bad.mkdir()
    x2 = _img()
    runner2 = OptimizationRunner(TinyModel(), x2, _cfg(log_loss=str(bad)),
                                 optimizer=torch.optim.Adam([x2], lr=0.1), progress_bar=Bar(),
                                 callbacks=OptimizationCallbacks(on_logging_error=errors.append))
    _, hist2, _ = runner2.run()
    assert len(errors) == 1 and isinstance(errors[0], OSError) and len(hist2["total_loss"]) == 6


def test_runner_argument_and_default_optimizer_rules() -> None:
    x = _img()
    with pytest.raises(ValueError, match="either optimizer or optimizer_factory"):
        OptimizationRunner(TinyModel(), x, _cfg(), optimizer=torch.optim.Adam([x]),
                           optimizer_factory=lambda p: torch.optim.Adam([p]))
    cfg = _cfg()
    cfg.optimization.lr = 0.5
    cfg.optimization.lbfgs_max_iter = 3
    r = OptimizationRunner(TinyModel(), x, cfg, progress_bar=Bar())
    assert isinstance(r.optimizer, torch.optim.LBFGS)
    assert r.optimizer.param_groups[0]["lr"] == 0.5 and r.optimizer.param_groups[0]["max_iter"] == 3
    made = []
    r2 = OptimizationRunner(TinyModel(), x, cfg, progress_bar=Bar(),
                            optimizer_factory=lambda p: made.append(p) or torch.optim.SGD([p], lr=0.1))
    assert made[0] is x and isinstance(r2.optimizer, torch.optim.SGD)
    with pytest.raises(RuntimeError, match="Progress bar not initialized"):
        _ = OptimizationRunner(TinyModel(), x, cfg, optimizer=torch.optim.SGD([x], lr=0.1)).progress_bar


def test_runner_closure_after_completion_and_missing_metrics() -> None:
    x = _img()
    runner = OptimizationRunner(TinyModel(), x, _cfg(steps=2),
                                optimizer=torch.optim.Adam([x], lr=0.1), progress_bar=Bar())
    assert float(runner._final_loss_tensor()) == 0.0  # noqa: SLF001
    runner.run()
    calls = runner._closure_calls  # noqa: SLF001
    again = runner._closure()  # noqa: SLF001
    assert runner._closure_calls == calls + 1  # noqa: SLF001
    assert float(again) == pytest.approx(float(runner._last_loss_tensor))  # noqa: SLF001

    class NoClosure(torch.optim.SGD):
        def step(self, closure=None):  # noqa: ANN001, ANN201
            return None

    x2 = _img()
    bad = OptimizationRunner(TinyModel(), x2, _cfg(steps=1), optimizer=NoClosure([x2], lr=0.1),
                             progress_bar=Bar())
    with pytest.raises(RuntimeError, match="did not record metrics for step 1"):
        bad.run()


def test_runner_non_finite_warnings(caplog: pytest.LogCaptureFixture) -> None:
    from style_transfer_visualizer_b200.logging_utils import logger

    class NanModel(nn.Module):
        def forward(self, x):  # noqa: ANN001, ANN201
            return [x.mean() * float("nan")], [x.mean()]

    x = _img()
    runner = OptimizationRunner(NanModel(), x, _cfg(steps=1),
                                optimizer=torch.optim.SGD([x], lr=0.0), progress_bar=Bar())
    logger.propagate = True
    try:
        with caplog.at_level(logging.WARNING, logger="style_transfer"):
            runner.run()
    finally:
        logger.propagate = False
    assert "Non-finite style score at step 1" in caplog.text
    assert "Non-finite total loss at step 1, using previous loss" in caplog.text
    assert "Non-finite content score" not in caplog.text


def test_runner_intro_crossfade_once() -> None:
    x = _img()
    sink = Sink()
    intro = np.zeros((8, 8, 3), dtype=np.uint8)
    runner = OptimizationRunner(TinyModel(), x, _cfg(steps=4, save_every=2),
                                optimizer=torch.optim.Adam([x], lr=0.1), progress_bar=Bar(),
                                video_writer=sink, intro_last_frame=intro, intro_crossfade_frames=3)
    runner.run()
    assert len(sink.frames) == 3 + 2 and runner.intro_transition_done
    assert runner.intro_last_frame is None
    blend = sink.frames[1].astype(int)
    assert np.all(blend <= sink.frames[3].astype(int))  # halfway between black and frame 1


def test_history_is_bounded(caplog: pytest.LogCaptureFixture) -> None:
    x = torch.rand(1, 3, 2, 2).requires_grad_(True)
    cfg = _cfg(steps=2100, save_every=5000, log_every=500)
    runner = OptimizationRunner(TinyModel(), x, cfg, optimizer=torch.optim.SGD([x], lr=0.01),
                                progress_bar=Bar())
    _, hist, _ = runner.run()
    assert len(hist["total_loss"]) == 2048


def test_image_load_path_matches_reference_semantics(tmp_path, caplog) -> None:  # noqa: ANN001
    """reference image_io.py:24-115 -- names, errors, warning and (on CPU) the exact tensor that
    torchvision's ToTensor + Normalize produce."""
    import logging

    import numpy as np
    from PIL import Image
    from torchvision import transforms

    from style_transfer_visualizer_b200 import image_io
    from style_transfer_visualizer_b200.constants import IMAGENET_MEAN, IMAGENET_STD

    with pytest.raises(FileNotFoundError, match="Image file not found: "):
        image_io.load_image(str(tmp_path / "missing.png"))
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not an image")
    with pytest.raises(OSError, match="Error loading image"):
        image_io.load_image(str(bad))
    rng = np.random.default_rng(0)
    small = tmp_path / "small.png"
    Image.fromarray(rng.integers(0, 256, (32, 80, 3), dtype=np.uint8)).save(small)
    with pytest.raises(ValueError, match=r"Image too small: 80x32\. Minimum dimension is 64px\."):
        image_io.load_image_to_tensor(str(small), torch.device("cpu"))
    ok = tmp_path / "ok.png"
    arr = rng.integers(0, 256, (70, 94, 3), dtype=np.uint8)
    Image.fromarray(arr).save(ok)
    cpu = torch.device("cpu")
    for normalize in (False, True):
        got = image_io.load_image_to_tensor(str(ok), cpu, normalize=normalize)
        pipeline = [transforms.ToTensor()]
        if normalize:
            pipeline.append(transforms.Normalize(mean=IMAGENET_MEAN, std=IMAGENET_STD))
        want = transforms.Compose(pipeline)(Image.open(ok).convert("RGB")).unsqueeze(0)
        assert got.shape == (1, 3, 70, 94) and got.dtype == torch.float32
        assert torch.equal(got, want)
    grey = tmp_path / "grey.png"
    Image.fromarray(arr[..., 0]).save(grey)   # converted to RGB on load
    assert image_io.load_image(str(grey)).mode == "RGB"

    class Big:
        width, height = 4000, 100

    caplog.set_level(logging.WARNING)
    image_io.validate_image_dimensions(Big())
    assert any("Image is large: 4000x100" in r.getMessage() for r in caplog.records)


def test_runner_gif_collector_and_postfix(caplog) -> None:  # noqa: ANN001
    """reference tests/test_optimization.py:394-441 (GIF frames mirror the video frames, the intro
    crossfade goes to the GIF only when gif_include_intro), :1165-1223 (postfix from the latest
    logged metrics), :725-742 (no summary line when nothing ran)."""
    import logging

    x = _img()
    video, gif, bar = Sink(), Sink(), Bar()
    cfg = _cfg(steps=4, save_every=2, log_every=2)
    cfg.video.gif_include_intro = True
    intro = np.zeros((8, 8, 3), dtype=np.uint8)
    runner = OptimizationRunner(TinyModel(), x, cfg, optimizer=torch.optim.Adam([x], lr=0.1),
                                progress_bar=bar, video_writer=video, gif_collector=gif,
                                intro_last_frame=intro, intro_crossfade_frames=3)
    runner.run()
    # two timelapse frames each; the crossfade (3 blended frames) precedes the first one in both
    assert len(video.frames) == 3 + 2 and len(gif.frames) == 3 + 2
    assert np.array_equal(video.frames[-1], gif.frames[-1])
    assert runner.intro_transition_done and runner.intro_last_frame is None
    assert bar.postfix and set(bar.postfix[-1]) == {"style", "content", "loss"}
    assert all(isinstance(v, str) for v in bar.postfix[-1].values())

    x2 = _img()
    gif2 = Sink()
    cfg2 = _cfg(steps=2, save_every=1)
    cfg2.video.gif_include_intro = False
    OptimizationRunner(TinyModel(), x2, cfg2, optimizer=torch.optim.Adam([x2], lr=0.1),
                       progress_bar=Bar(), gif_collector=gif2, intro_last_frame=intro,
                       intro_crossfade_frames=3).run()
    assert len(gif2.frames) == 2  # no crossfade in the GIF

    caplog.set_level(logging.INFO)
    caplog.clear()
    x3 = _img()
    idle = OptimizationRunner(TinyModel(), x3, _cfg(steps=2), optimizer=torch.optim.Adam([x3]),
                              progress_bar=Bar())
    idle._log_optimization_summary()  # noqa: SLF001
    assert not any("Optimization finished" in r.getMessage() for r in caplog.records)


def test_closure_reads_no_scalars_between_flushes(monkeypatch) -> None:  # noqa: ANN001
    """reference tests/test_optimization.py:943-970: between log_every flushes the closure must not
    pull scalars to the host (`.item()` is a device sync on CUDA)."""
    calls = {"n": 0}
    orig = torch.Tensor.item

    def counting_item(self):  # noqa: ANN001, ANN202
        calls["n"] += 1
        return orig(self)

    x = _img()
    runner = OptimizationRunner(TinyModel(), x, _cfg(steps=9, log_every=5, save_every=100),
                                optimizer=torch.optim.SGD([x], lr=0.01), progress_bar=Bar())
    monkeypatch.setattr(torch.Tensor, "item", counting_item)
    runner.run()
    monkeypatch.undo()
    # one flush at step 5 (three scalars) -- nothing at steps 1-4 and 6-9
    assert calls["n"] <= 3


# --------------------------------------------------------------------------- round-2 host logic
def test_accumulator_adopts_rows_written_by_the_step_graph() -> None:
    """``adopt_device_rows``: the captured step graph writes row ``k % capacity`` itself
    (``stv_step_scores``); ``accumulate`` then only keeps the books, ``export_history`` returns the
    retained window oldest-first and the logged scalars come from the row, not from the tensors."""
    from style_transfer_visualizer_b200.loss_accumulator import LossAccumulator

    cap = 4
    acc = LossAccumulator(log_every=2, history_capacity=cap, track_history=True,
                          device=torch.device("cpu"), dtype=torch.float32)
    rows = torch.zeros(cap, 3)
    acc.adopt_device_rows(rows)
    junk = torch.tensor(-1.0)
    logged = []
    for step in range(1, 7):                        # the "graph" writes the row, then accumulate()
        rows[(step - 1) % cap] = torch.tensor([step, 10.0 * step, 100.0 * step])
        out = acc.accumulate(step, junk, junk, junk)
        if out is not None:
            logged.append((out.step, out.style_loss, out.content_loss, out.total_loss))
    assert logged == [(2, 2.0, 20.0, 200.0), (4, 4.0, 40.0, 400.0), (6, 6.0, 60.0, 600.0)]
    hist = acc.export_history()
    assert hist["style_loss"] == [3.0, 4.0, 5.0, 6.0]      # the last `cap` steps, oldest first
    assert hist["total_loss"] == [300.0, 400.0, 500.0, 600.0]
    assert acc.history_truncated
    with pytest.raises(RuntimeError, match="must precede"):
        acc.adopt_device_rows(rows)
    with pytest.raises(ValueError, match="shape"):
        LossAccumulator(log_every=1, history_capacity=cap, track_history=True,
                        device=torch.device("cpu"), dtype=torch.float32
                        ).adopt_device_rows(torch.zeros(cap + 1, 3))


def test_bench_accounting_helpers() -> None:
    """bench.py: every tensor-core conv entry point is counted with its algorithmic FLOPs, and the
    TF32 roofline denominator follows the clock / power regime sampled during the timed blocks."""
    import bench

    h, w, cin, cout = 8, 16, 64, 128
    fwd = 2.0 * 9 * h * w * cin * cout
    for name in bench.CONV_FWD:
        assert bench.conv_flops(name, (0, 0, 0, h, w, cin, cout)) == fwd
    for name in bench.CONV_DGRAD:
        assert bench.conv_flops(name, (0, 0, h, w, cout, cin)) == fwd
    assert bench.conv_flops("stv_conv3x3_first_dgrad_tc", (0, 0, h, w, 64)) == 2.0 * 9 * h * w * 64 * 3
    assert bench.conv_flops("stv_conv3x3_first_dgrad_rows", (0, 0, h, w, 64)) == 2.0 * 9 * h * w * 64 * 3
    assert bench.conv_flops("stv_style_bwd", (0, 0, 100, 64)) == 2.0 * 100 * 64 * 64
    assert bench.conv_flops("stv_adam_step", ()) is None
    # every conv entry point the native layer exports is covered by the accounting
    conv_like = {n for n in nat.EXPORTED_SYMBOLS
                 if n.startswith(("stv_conv3x3_fwd", "stv_conv3x3_dgrad"))}
    assert conv_like <= set(bench.CONV_FWD) | set(bench.CONV_DGRAD) | {"stv_conv3x3_dgrad_bits_style"}
    assert bench.conv_flops("stv_conv3x3_dgrad_bits_style", (0, 0, h, w, cout, cin)) == \
        fwd + 2.0 * h * w * cin * cin
    peaks = {"bf16_burst": 1646.0, "bf16_sustained": 1399.4, "hbm_gbs": 6546.2, "source": "x"}
    full = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []}
    capped = {"sm_mhz": 1740.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"]}
    assert bench.tensor_peak(peaks, full)[0] == 823.0
    assert bench.tensor_peak(peaks, capped)[0] == 699.7
    assert bench.tensor_peak(peaks, None)[0] == 823.0


def test_band_rows_and_arena_layout() -> None:
    from style_transfer_visualizer_b200.sharded import plan_bands, rows_at

    assert [rows_at(272, k) for k in range(5)] == [272, 136, 68, 34, 17]
    assert rows_at(135, 1) == 67 and rows_at(135, 2) == 33      # floor mode, level by level
    bands = plan_bands(2160, 8)
    assert [b - a for a, b in bands] == [272] * 7 + [256]
    assert all(a % 16 == 0 for a, _ in bands) and bands[-1][1] == 2160


def test_first_layer_dgrad_rows_formulation_matches_autograd() -> None:
    """The identity behind ``conv_first_dgrad_tc.cu`` (x taps folded into the GEMM's N), in fp64
    on the CPU: with ``w_rows[t][kx*3+ci][co] = w[co][ci][2-t][kx]`` (include/stv_b200.h),
    ``U[y][x'][kx][ci] = sum_{t,co} dY[y-1+t][x'][co] * w_rows[t][kx*3+ci][co]`` and
    ``dimg[y][x][ci] = U[y][x+1][0][ci] + U[y][x][1][ci] + U[y][x-1][2][ci]`` equal the input
    gradient of the reference's first Conv2d (zero padding at the borders)."""
    import numpy as np
    import torch

    rng = np.random.default_rng(5)
    h, w_, co = 7, 9, 64
    wt = rng.standard_normal((co, 3, 3, 3))
    dy = rng.standard_normal((1, co, h, w_))
    ref = torch.nn.grad.conv2d_input((1, 3, h, w_), torch.from_numpy(wt), torch.from_numpy(dy),
                                     padding=1)[0].numpy()                      # [3, H, W]
    w_rows = np.zeros((3, 16, co))
    for t in range(3):
        for kx in range(3):
            for ci in range(3):
                w_rows[t, kx * 3 + ci] = wt[:, ci, 2 - t, kx]
    dyp = np.zeros((h + 2, w_ + 2, co))
    dyp[1:-1, 1:-1] = dy[0].transpose(1, 2, 0)                                  # NHWC, zero halo
    u = np.zeros((h, w_ + 2, 16))                                               # x' = -1 .. W
    for t in range(3):
        u += np.einsum("yxc,nc->yxn", dyp[t:t + h], w_rows[t])
    out = np.zeros((3, h, w_))
    for ci in range(3):
        out[ci] = u[:, 2:, 0 + ci] + u[:, 1:-1, 3 + ci] + u[:, :-2, 6 + ci]
    assert np.allclose(out, ref, rtol=1e-10, atol=1e-10)
    assert not w_rows[:, 9:].any()
