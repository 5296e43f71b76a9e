"""Whole-step CUDA graph: forward, losses, weighted total, backward and the Adam update replayed
with a single launch.

At 512x512 one step is ~60 kernels of a few microseconds to ~100 us each; issued one by one the
step is bound by launch overhead and by the host syncs the reference performs every closure
("512^2 is launch/sync-bound").  Capturing the fixed kernel program once and
replaying it removes both.  The step counter of Adam's bias correction lives on the device
(``stv_adam_step_dev``), so the captured graph needs no per-step host values.
"""
from __future__ import annotations

import torch

from . import _native as nat
from . import ops
from .core_model import StyleContentModel
from .optim import FusedAdam, FusedLBFGS


class FusedStep:
    """Graph-captured optimisation step for (StyleContentModel, FusedAdam | FusedLBFGS[max_iter=1])."""

    @classmethod
    def try_create(cls, model: object, x: torch.Tensor, optimizer: object, style_w: float,
                   content_w: float, *, record_capacity: int = 0) -> "FusedStep | None":
        if not isinstance(model, StyleContentModel):
            return None
        is_lbfgs = isinstance(optimizer, FusedLBFGS) and \
            optimizer.param_groups[0]["max_iter"] == 1
        if not isinstance(optimizer, FusedAdam) and not is_lbfgs:
            return None
        if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
            return None
        if model.style_targets is None or model.content_targets is None:
            return None
        params = [p for g in optimizer.param_groups for p in g["params"]]
        if len(params) != 1 or params[0] is not x:
            return None
        return cls(model, x, optimizer, style_w, content_w, record_capacity=record_capacity)

    def __init__(self, model: StyleContentModel, x: torch.Tensor,
                 optimizer: "FusedAdam | FusedLBFGS", style_w: float, content_w: float, *,
                 record_capacity: int = 0) -> None:
        self.x = x
        self.optimizer = optimizer
        self.engine = model.engine_for(x.device)
        if self.engine.style_targets is None:
            msg = "model targets are not set on this device"
            raise RuntimeError(msg)
        self.style_w = float(style_w)
        self.content_w = float(content_w)
        self.n_style = len(self.engine.style_idx)
        self.n_content = len(self.engine.content_idx)
        dev = x.device
        self.grad_w = torch.tensor([self.style_w] * self.n_style
                                   + [self.content_w] * self.n_content,
                                   device=dev, dtype=torch.float32)
        group = optimizer.param_groups[0]
        self.lr = float(group["lr"])
        self.is_adam = isinstance(optimizer, FusedAdam)
        state = optimizer.state[x]
        if self.is_adam:
            self.beta1, self.beta2 = (float(b) for b in group["betas"])
            self.eps = float(group["eps"])
            if not state:
                state["step"] = 0
                state["exp_avg"] = torch.zeros_like(x)
                state["exp_avg_sq"] = torch.zeros_like(x)
            self.adam_state = torch.zeros(3, device=dev, dtype=torch.float32)
            self.adam_state[0] = float(state["step"])
        else:
            optimizer._device_state()  # noqa: SLF001  (history buffers allocated before capture)
        self.state = state
        self.scores = torch.zeros(3, device=dev, dtype=torch.float32)   # style, content, total
        # Per-step records written INSIDE the graph at a device-side step counter (reference
        # optimization.py:375-391 finiteness checks, :402-422 loss recording): the runner reads rows
        # at its logging cadence instead of launching bookkeeping kernels every step.
        self.record_capacity = int(record_capacity)
        self.loss_ring: torch.Tensor | None = None      # [capacity, 3]
        self.finite_ring: torch.Tensor | None = None    # [capacity] int32 bit flags
        self.record_counter: torch.Tensor | None = None  # [1] int32: rows written so far
        if self.record_capacity > 0:
            self.loss_ring = torch.zeros(self.record_capacity, 3, device=dev, dtype=torch.float32)
            self.finite_ring = torch.zeros(self.record_capacity, device=dev, dtype=torch.int32)
            self.record_counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self.height, self.width = int(x.shape[2]), int(x.shape[3])
        self.graph: torch.cuda.CUDAGraph | None = None
        self.kernel_launches = 0

    def _forward_backward(self) -> torch.Tensor:
        losses, _gen = self.engine.forward_losses(self.x.detach())
        self._losses = losses
        return self.engine.backward_losses(self.height, self.width, self.grad_w)

    def _body(self) -> None:
        grad = self._forward_backward()
        # weighted total, finiteness flags and the history row: one single-thread kernel
        ops.step_scores(self._losses, self.n_style, self.n_content, self.style_w, self.content_w,
                        self.scores, loss_ring=self.loss_ring, finite_ring=self.finite_ring,
                        counter=self.record_counter)
        if self.is_adam:
            ops.adam_step_dev(self.x.detach(), grad, self.state["exp_avg"],
                              self.state["exp_avg_sq"], self.adam_state, lr=self.lr,
                              beta1=self.beta1, beta2=self.beta2, eps=self.eps)
        else:
            self.optimizer.device_step(grad)

    def _capture(self) -> None:
        with torch.no_grad():
            # dry run without the update: allocates workspaces, sets kernel attributes -- only the
            # first time this engine sees the image size (a second run on the same model, e.g. the
            # next job of the same shape, captures straight away)
            if not self.engine.is_warm(self.height, self.width):
                side = torch.cuda.Stream(device=self.x.device)
                side.wait_stream(torch.cuda.current_stream(self.x.device))
                with torch.cuda.stream(side):
                    self._forward_backward()
                torch.cuda.current_stream(self.x.device).wait_stream(side)
            torch.cuda.synchronize(self.x.device)
            graph = torch.cuda.CUDAGraph()
            before = nat.launch_count()
            with torch.cuda.graph(graph):
                self._body()
            self.kernel_launches = nat.launch_count() - before  # this library's kernels per replay
            self.graph = graph
        # like the reference, the image carries the last gradient after each step
        self.x.grad = self.engine.grad_buffer(self.height, self.width)

    def step(self) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One optimisation step; returns device scalars (style, content, total) that are
        overwritten by the next call."""
        if self.graph is None:
            self._capture()
        self.graph.replay()
        if self.is_adam:
            self.state["step"] += 1
        else:
            self.state["func_evals"] = self.state.get("func_evals", 0) + 1
        return self.scores[0], self.scores[1], self.scores[2]
