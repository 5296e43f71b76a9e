"""Optimisation loop with the reference's runner surface (optimization.py:32-529).

``OptimizationRunner`` keeps the reference's constructor, ``run()`` contract, closure semantics
(one accepted step = one ``optimizer.step(closure)``; frame / progress / callbacks once per
accepted step even if the optimiser evaluates the closure several times), loss recording through
``LossAccumulator`` / ``LossCSVLogger``, timelapse frame emission and callbacks.

What differs is how the device is driven:
  * finiteness checks (reference: three blocking ``if not torch.isfinite(t)`` syncs per closure)
    are recorded on the device and turned into the same warnings at the logging cadence;
  * timelapse frames go through ``FrameReadback`` (fused u8 kernel + pinned async D2H on a side
    stream) and are handed to the sinks in order, one emission late, so the compute stream never
    stalls on the copy or the encoder;
  * when the model is this package's ``StyleContentModel`` and the optimiser its ``FusedAdam``,
    the whole step (forward, losses, backward, update) is replayed from one CUDA graph.
"""
from __future__ import annotations

import logging
import time
from collections.abc import Callable, Mapping
from dataclasses import dataclass
from typing import Protocol

import numpy as np
import torch
from torch import nn
from torch.optim import Optimizer

from . import image_io as stv_image_io
from . import video as stv_video
from .config import StyleTransferConfig
from .constants import CSV_LOGGING_RECOMMENDED_STEPS, LossHistory
from .logging_utils import logger
from .loss_accumulator import DEFAULT_HISTORY_CAPACITY, LoggedLoss, LossAccumulator
from .loss_logger import LossCSVLogger


class ProgressReporter(Protocol):
    """The subset of tqdm's interface the runner relies on."""

    def update(self, n: float | None = 1) -> bool | None: ...

    def set_postfix(self, ordered_dict: Mapping[str, object] | None = None,
                    refresh: bool | None = True, **kwargs: object) -> None: ...  # noqa: FBT001, FBT002

    def close(self) -> None: ...


@dataclass(slots=True)
class StepMetrics:
    """Host-synced scalar losses exposed to callbacks (populated on ``log_every`` steps only)."""

    step: int
    style_loss: float | None = None
    content_loss: float | None = None
    total_loss: float | None = None

    @property
    def has_values(self) -> bool:
        return None not in (self.style_loss, self.content_loss, self.total_loss)


@dataclass(slots=True)
class StepTensors:
    """Per-step loss tensors kept on the device."""

    step: int
    style_score: torch.Tensor
    content_score: torch.Tensor
    total_loss: torch.Tensor


@dataclass(slots=True)
class OptimizationCallbacks:
    """Optional hooks invoked around optimisation events."""

    on_step_start: Callable[[int], None] | None = None
    on_step_end: Callable[[StepMetrics], None] | None = None
    on_video_frame: Callable[[np.ndarray, int], None] | None = None
    on_logging_error: Callable[[Exception], None] | None = None


_FINITE_NAMES = ("style score", "content score", "total loss")


class OptimizationRunner:
    """Drive the optimisation loop, logging, frame emission and progress reporting."""

    def __init__(  # noqa: PLR0913
        self,
        model: nn.Module,
        input_img: torch.Tensor,
        config: StyleTransferConfig,
        *,
        optimizer: Optimizer | None = None,
        optimizer_factory: Callable[[torch.Tensor], Optimizer] | None = None,
        progress_bar: ProgressReporter | None = None,
        callbacks: OptimizationCallbacks | None = None,
        video_writer: stv_video.VideoFrameSink | None = None,
        gif_collector: stv_video.VideoFrameSink | None = None,
        intro_last_frame: np.ndarray | None = None,
        intro_crossfade_frames: int = 0,
        use_cuda_graph: bool | None = None,
        async_frames: bool | None = None,
    ) -> None:
        if optimizer is not None and optimizer_factory is not None:
            msg = "Provide either optimizer or optimizer_factory, not both."
            raise ValueError(msg)

        self.model = model
        self.input_img = input_img
        self.config = config
        self.optimizer = optimizer if optimizer is not None \
            else self._build_optimizer(optimizer_factory)

        self._progress_bar: ProgressReporter | None = progress_bar
        self._owns_progress_bar = False
        self.callbacks = callbacks or OptimizationCallbacks()

        self.video_writer = video_writer
        self.gif_collector = gif_collector
        self.intro_last_frame = intro_last_frame
        self.intro_crossfade_frames = intro_crossfade_frames
        self.intro_transition_done = intro_last_frame is None

        self.loss_logger: LossCSVLogger | None = None
        self._loss_accumulator: LossAccumulator | None = None
        self._latest_logged: LoggedLoss | None = None
        self._last_loss_tensor: torch.Tensor | None = None
        self._configure_logging()

        self._step_index = 0
        self._active_step_idx: int | None = None
        self._pending_step_tensors: StepTensors | None = None
        self._closure_calls = 0

        on_cuda = input_img.is_cuda
        self._async_frames = on_cuda if async_frames is None else (async_frames and on_cuda)
        self._readback: stv_image_io.FrameReadback | None = None
        self._lazy_finite = on_cuda
        self._finite_ring: torch.Tensor | None = None     # [window] int32 bit flags (eager steps)
        self._finite_vals: torch.Tensor | None = None     # [3] staging for stv_finite_flags
        self._finite_steps: list[int] = []
        self._graph_records = False   # finiteness + history rows written inside the step graph
        self._graph_rows_written = 0
        self._use_cuda_graph = use_cuda_graph
        self._fused = None  # FusedStep, built lazily in run()
        self.frames_emitted = 0
        self.frame_bytes_d2h = 0

    # ------------------------------------------------------------------ properties
    @property
    def progress_bar(self) -> ProgressReporter:
        if self._progress_bar is None:
            msg = "Progress bar not initialized. Call run() before use."
            raise RuntimeError(msg)
        return self._progress_bar

    @property
    def total_steps(self) -> int:
        return self.config.optimization.steps

    # ------------------------------------------------------------------ main loop
    def run(self) -> tuple[torch.Tensor, LossHistory, float]:
        """Execute all steps; returns ``(input_img, loss history, elapsed seconds)``."""
        self.prepare()
        start = time.time()
        try:
            while self._step_index < self.total_steps:
                step_idx = self._step_index + 1
                self._emit_step_start(step_idx)
                self._active_step_idx = step_idx
                self._pending_step_tensors = None
                try:
                    if self._fused is not None:
                        self._fused_step(step_idx)
                    else:
                        self.optimizer.step(self._closure)  # type: ignore[arg-type]
                finally:
                    self._active_step_idx = None
                tensors = self._pending_step_tensors
                if tensors is None:
                    msg = f"Optimizer closure did not record metrics for step {step_idx}"
                    raise RuntimeError(msg)
                self._finalize_step(tensors)
                self._pending_step_tensors = None
            self._drain_frames()
            self._flush_finite_checks()
        except BaseException:
            # A failing or interrupted run still owes the sinks the frames already submitted and the
            # log the deferred "Non-finite ..." warnings -- on CUDA often the only explanation of the
            # failure (the reference emitted them synchronously).  Best effort: never mask the error.
            for late in (self._drain_frames, self._flush_finite_checks):
                try:
                    late()
                except Exception:  # noqa: BLE001, S110
                    pass
            raise
        finally:
            self._cleanup()

        elapsed = time.time() - start
        self._log_optimization_summary()
        acc = self._loss_accumulator
        history: LossHistory = acc.export_history() if acc is not None and acc.tracks_history \
            else {}
        return self.input_img, history, elapsed

    def prepare(self) -> None:
        """One-off set-up hoisted out of the step loop: progress bar, whole-step CUDA graph and
        the pinned frame-readback ring (pinned allocations take milliseconds)."""
        self._ensure_progress_bar()
        self._maybe_build_fused_step()
        wants_frames = (self.video_writer is not None or self.gif_collector is not None) and \
            self.config.video.save_every <= self.total_steps
        if self._async_frames and wants_frames and self._readback is None:
            h, w = int(self.input_img.shape[-2]), int(self.input_img.shape[-1])
            self._readback = stv_image_io.shared_readback(self.input_img.device, h, w)

    def _build_optimizer(
        self, optimizer_factory: Callable[[torch.Tensor], Optimizer] | None,
    ) -> Optimizer:
        """Default optimiser: L-BFGS with the configured lr / max_iter / max_eval (reference
        optimization.py:204-217), on this package's kernels when the image lives on the GPU."""
        if optimizer_factory is not None:
            return optimizer_factory(self.input_img)
        opt = self.config.optimization
        if self.input_img.is_cuda:
            from .optim import FusedLBFGS

            return FusedLBFGS([self.input_img], lr=opt.lr, max_iter=opt.lbfgs_max_iter,
                              max_eval=opt.lbfgs_max_eval)
        return torch.optim.LBFGS([self.input_img], lr=opt.lr, max_iter=opt.lbfgs_max_iter,
                                 max_eval=opt.lbfgs_max_eval)

    def _configure_logging(self) -> None:
        """CSV logging when requested (falling back to in-memory history on I/O errors), plus the
        bounded device-side history (reference optimization.py:219-263)."""
        out = self.config.output
        steps = self.total_steps
        track_history = True
        self.loss_logger = None
        if out.log_loss:
            try:
                self.loss_logger = LossCSVLogger(out.log_loss, out.log_every)
            except OSError as exc:
                logger.error("Failed to initialize CSV logging: %s", exc)
                if self.callbacks.on_logging_error is not None:
                    self.callbacks.on_logging_error(exc)
            else:
                logger.info("Loss CSV logging enabled: %s", out.log_loss)
                track_history = False

        capacity = min(steps, DEFAULT_HISTORY_CAPACITY)
        self._loss_accumulator = LossAccumulator(
            log_every=out.log_every, history_capacity=capacity, track_history=track_history,
            device=self.input_img.device, dtype=self.input_img.dtype)

        if track_history and steps > capacity:
            logger.warning(
                "Long run detected (%d steps). In-memory loss history is capped at %d entries; "
                "enable --log-loss for a full CSV.", steps, capacity)
        elif track_history and steps > CSV_LOGGING_RECOMMENDED_STEPS:
            logger.warning(
                "Long run detected (%d steps). Consider enabling --log-loss to capture every "
                "step.", steps)

    def _ensure_progress_bar(self) -> None:
        if self._progress_bar is None:
            from tqdm import tqdm

            self._progress_bar = tqdm(total=self.total_steps, desc="Style Transfer")
            self._owns_progress_bar = True

    # ------------------------------------------------------------------ one step
    def _closure(self) -> torch.Tensor:
        """Closure handed to the optimiser (may be evaluated several times per step)."""
        self._closure_calls += 1
        if self._step_index >= self.total_steps:
            return self._final_loss_tensor()
        step_idx = self._active_step_idx or (self._step_index + 1)
        tensors = self._run_single_step(step_idx)
        self._pending_step_tensors = tensors
        return tensors.total_loss

    def _run_single_step(self, step_idx: int) -> StepTensors:
        """zero_grad -> losses -> weighted total -> backward (reference optimization.py:286-327)."""
        opt_cfg = self.config.optimization
        self.optimizer.zero_grad()
        style_losses, content_losses = self.model(self.input_img)

        zero = torch.zeros((), device=self.input_img.device, dtype=self.input_img.dtype)
        style_score = torch.stack(style_losses).sum() if style_losses else zero
        content_score = torch.stack(content_losses).sum() if content_losses else zero
        loss = opt_cfg.style_w * style_score + opt_cfg.content_w * content_score
        loss.backward()

        self._check_finite(style_score, content_score, loss, step_idx)
        return StepTensors(step=step_idx, style_score=style_score, content_score=content_score,
                           total_loss=loss)

    def _finalize_step(self, tensors: StepTensors) -> None:
        """Bookkeeping after an accepted optimiser step."""
        self._step_index = tensors.step
        self._last_loss_tensor = tensors.total_loss.detach()

        logged = self._record_losses(tensors)
        if logged is None:
            metrics = StepMetrics(step=tensors.step)
        else:
            self._latest_logged = logged
            metrics = StepMetrics(step=logged.step, style_loss=logged.style_loss,
                                  content_loss=logged.content_loss, total_loss=logged.total_loss)
            if self._graph_records and self._finite_steps == [logged.step]:
                # the only pending step is the one whose three losses were just read back: its
                # flags follow from those host values -- no second device-to-host copy (with
                # log_every=1 that copy was a second sync in every step)
                import math

                self._finite_steps.clear()
                for which, val in enumerate((logged.style_loss, logged.content_loss,
                                             logged.total_loss)):
                    if not math.isfinite(val):
                        self._warn_non_finite(which, logged.step)
            else:
                self._flush_finite_checks()  # the host has just synced anyway

        self._maybe_write_video_frame(metrics)
        self.progress_bar.update(1)
        self._emit_step_end(metrics)

    def _final_loss_tensor(self) -> torch.Tensor:
        if self._last_loss_tensor is not None:
            return self._last_loss_tensor.detach()
        return torch.zeros((), device=self.input_img.device, dtype=self.input_img.dtype)

    def _log_optimization_summary(self) -> None:
        if self._step_index <= 0:
            return
        logger.info(
            "Optimization finished with %d accepted steps and %d closure evaluations "
            "(%.2f closures/step).", self._step_index, self._closure_calls,
            self._closure_calls / self._step_index)

    # ------------------------------------------------------------------ finiteness
    @staticmethod
    def _warn_non_finite(which: int, step_idx: int) -> None:
        if which == 0:
            logger.warning("Non-finite style score at step %d", step_idx)
        elif which == 1:
            logger.warning("Non-finite content score at step %d", step_idx)
        else:
            logger.warning("Non-finite total loss at step %d, using previous loss", step_idx)

    def _check_finite(self, style_score: torch.Tensor, content_score: torch.Tensor,
                      total_loss: torch.Tensor, step_idx: int) -> None:
        """Warn when a recorded loss is non-finite (same messages as the reference,
        optimization.py:375-400).  On CUDA the three flags are written to a device ring and the
        warnings are emitted at the next logging sync instead of stalling every closure."""
        if self._graph_records:
            # the captured graph already wrote this step's flags (stv_step_scores) to its ring
            if len(self._finite_steps) == self._fused.record_capacity:
                self._flush_finite_checks()
            self._finite_steps.append(step_idx)
        elif self._lazy_finite and style_score.is_cuda:
            from . import ops

            if self._finite_ring is None:
                window = max(1, min(self.config.output.log_every, 1024))
                self._finite_ring = torch.zeros(window, 3, device=style_score.device,
                                                dtype=torch.int32)
                self._finite_vals = torch.zeros(3, device=style_score.device, dtype=torch.float32)
            if len(self._finite_steps) == self._finite_ring.shape[0]:
                self._flush_finite_checks()
            row = len(self._finite_steps)
            torch.stack((style_score.detach(), content_score.detach(), total_loss.detach()),
                        out=self._finite_vals)
            ops.finite_flags(self._finite_vals, self._finite_ring[row])  # flags |= !isfinite
            self._finite_steps.append(step_idx)
        else:
            for which, val in enumerate((style_score, content_score, total_loss)):
                if not torch.isfinite(val):
                    self._warn_non_finite(which, step_idx)

        if logger.isEnabledFor(logging.DEBUG):
            logger.debug("Step %d: Style %.4e, Content %.4e, Total %.4e", step_idx,
                         float(style_score.detach().item()), float(content_score.detach().item()),
                         float(total_loss.detach().item()))

    def _flush_finite_checks(self) -> None:
        """Read the pending finiteness flags (one small D2H) and emit the warnings in order."""
        if not self._finite_steps:
            return
        n = len(self._finite_steps)
        if self._graph_records:
            fused = self._fused
            cap = fused.record_capacity
            ring = fused.finite_ring.cpu()
            # the graph has written rows 0 .. self._step_records - 1 (mod capacity); the pending
            # steps are the last n of them
            last = self._graph_rows_written
            rows = [int(ring[(last - n + i) % cap]) for i in range(n)]
            flags = [[bool(r & 1), bool(r & 2), bool(r & 4)] for r in rows]
        elif self._finite_ring is not None:
            host = self._finite_ring[:n].cpu()
            flags = [[bool(host[r, w]) for w in range(3)] for r in range(n)]
            self._finite_ring.zero_()
        else:
            return
        for row, step_idx in enumerate(self._finite_steps):
            for which in range(3):
                if flags[row][which]:
                    self._warn_non_finite(which, step_idx)
        self._finite_steps.clear()

    # ------------------------------------------------------------------ losses
    def _record_losses(self, tensors: StepTensors) -> LoggedLoss | None:
        if self._loss_accumulator is None:
            return None
        logged = self._loss_accumulator.accumulate(
            tensors.step, tensors.style_score, tensors.content_score, tensors.total_loss)
        if logged is not None and self.loss_logger is not None:
            self.loss_logger.log(logged.step, logged.style_loss, logged.content_loss,
                                 logged.total_loss)
        return logged

    # ------------------------------------------------------------------ timelapse frames
    def _frame_due(self, step_idx: int) -> bool:
        save_every = self.config.video.save_every
        return bool(save_every) and step_idx % save_every == 0 and \
            (self.video_writer is not None or self.gif_collector is not None)

    def _maybe_write_video_frame(self, metrics: StepMetrics) -> None:
        """Emit a timelapse frame every ``save_every`` steps (reference optimization.py:424-489)."""
        step_idx = metrics.step
        if not self._frame_due(step_idx):
            return
        normalize = self.config.optimization.normalize
        if self._async_frames:
            if self._readback is None:
                h, w = int(self.input_img.shape[-2]), int(self.input_img.shape[-1])
                self._readback = stv_image_io.shared_readback(self.input_img.device, h, w)
            # hand over the previous frame (its copy finished long ago), then queue this one
            while self._readback.pending >= self._readback.depth - 1:
                self._deliver_frame(*self._readback.collect())
            self._readback.submit(self.input_img, normalize=normalize, tag=(step_idx, metrics))
            return
        with torch.no_grad():
            if self.input_img.is_cuda:
                img_np = stv_image_io.frame_to_numpy(self.input_img, normalize=normalize)
            else:
                image = stv_image_io.prepare_image_for_output(self.input_img,
                                                              normalize=normalize)
                if image is None:
                    return
                img_np = (image.squeeze(0).permute(1, 2, 0).cpu().numpy() * 255).astype("uint8")
        self._deliver_frame(img_np, (step_idx, metrics))

    def _drain_frames(self) -> None:
        if self._readback is not None:
            while self._readback.pending:
                self._deliver_frame(*self._readback.collect())

    def _deliver_frame(self, img_np: np.ndarray, tag: object) -> None:
        step_idx, metrics = tag  # type: ignore[misc]
        video_writer, gif_collector = self.video_writer, self.gif_collector
        self.frames_emitted += 1
        self.frame_bytes_d2h += img_np.nbytes

        if self.intro_last_frame is not None and not self.intro_transition_done:
            if video_writer is not None and self.config.video.intro_enabled:
                stv_video.append_crossfade(video_writer, self.intro_last_frame, img_np,
                                           self.intro_crossfade_frames)
            if gif_collector is not None and self.config.video.gif_include_intro:
                stv_video.append_crossfade(gif_collector, self.intro_last_frame, img_np,
                                           self.intro_crossfade_frames)
            self.intro_transition_done = True
            self.intro_last_frame = None

        if video_writer is not None:
            video_writer.append_data(img_np)
        if gif_collector is not None:
            gif_collector.append_data(img_np)
        self._update_progress_postfix(metrics)
        if self.callbacks.on_video_frame is not None:
            self.callbacks.on_video_frame(img_np, step_idx)

    # ------------------------------------------------------------------ hooks / display
    def _emit_step_start(self, step_idx: int) -> None:
        if self.callbacks.on_step_start is not None:
            self.callbacks.on_step_start(step_idx)

    def _emit_step_end(self, metrics: StepMetrics) -> None:
        if self.callbacks.on_step_end is not None:
            self.callbacks.on_step_end(metrics)

    def _update_progress_postfix(self, metrics: StepMetrics) -> None:
        shown = metrics if metrics.has_values else self._latest_logged
        if shown is None:
            return
        postfix = {}
        for label, value in (("style", shown.style_loss), ("content", shown.content_loss),
                             ("loss", shown.total_loss)):
            if value is not None:
                postfix[label] = f"{value:.4f}"
        if postfix:
            self.progress_bar.set_postfix(postfix)

    def _cleanup(self) -> None:
        if self.loss_logger is not None:
            self.loss_logger.close()
        if self._owns_progress_bar and self._progress_bar is not None:
            self._progress_bar.close()

    # ------------------------------------------------------------------ CUDA-graph fast path
    def _maybe_build_fused_step(self) -> None:
        """Use the whole-step CUDA graph when model and optimiser are this package's own."""
        if self._use_cuda_graph is False or self._fused is not None:
            return
        from .fused_step import FusedStep

        acc = self._loss_accumulator
        capacity = acc.capacity if acc is not None else min(self.total_steps,
                                                            DEFAULT_HISTORY_CAPACITY)
        self._fused = FusedStep.try_create(self.model, self.input_img, self.optimizer,
                                           self.config.optimization.style_w,
                                           self.config.optimization.content_w,
                                           record_capacity=capacity)
        if self._fused is not None and self._step_index == 0:
            # history rows and finiteness flags come from the graph itself from now on
            self._graph_records = True
            self._graph_rows_written = 0
            if acc is not None:
                acc.adopt_device_rows(self._fused.loss_ring)
        if self._fused is None and self._use_cuda_graph:
            msg = ("use_cuda_graph=True needs this package's StyleContentModel with targets set "
                   "and a FusedAdam or FusedLBFGS(max_iter=1) optimiser on the image")
            raise RuntimeError(msg)

    def _fused_step(self, step_idx: int) -> None:
        self._closure_calls += 1
        style_score, content_score, loss = self._fused.step()
        if self._graph_records:
            self._graph_rows_written += 1
        self._check_finite(style_score, content_score, loss, step_idx)
        self._pending_step_tensors = StepTensors(
            step=step_idx, style_score=style_score, content_score=content_score, total_loss=loss)
