"""Image-update optimisers on sm_100a kernels, as ``torch.optim.Optimizer`` subclasses so they
plug into the runner's ``optimizer=`` / ``optimizer_factory=`` seam (reference
optimization.py:104-105, 204-217) and ``prepare_model_and_input`` (core_model.py:344-349).

``FusedAdam`` follows torch.optim.Adam's single-tensor update (one fused pass instead of six
element-wise launches).  ``FusedLBFGS`` follows torch.optim.LBFGS's algorithm and defaults
(history 100, tolerance_grad 1e-7, tolerance_change 1e-9, no line search) including its
first-iteration step scaling and its early exits, with the vector work (dot / axpy / scale / abs
statistics) on the kernels in ``elementwise.cu``.
"""
from __future__ import annotations

from collections.abc import Callable, Iterable

import torch
from torch.optim import Optimizer

from . import _native as nat
from . import ops


def _single_param(optimizer: Optimizer) -> torch.Tensor:
    params = [p for group in optimizer.param_groups for p in group["params"]]
    if len(params) != 1:
        msg = f"{type(optimizer).__name__} optimises exactly one tensor (the image)"
        raise ValueError(msg)
    return params[0]


def _check_param(p: torch.Tensor) -> None:
    if p.dtype != torch.float32 or not p.is_contiguous():
        msg = "the optimised image must be a contiguous float32 tensor"
        raise ValueError(msg)
    nat.require_device(p.device)


class FusedAdam(Optimizer):
    """Adam (no weight decay / amsgrad) in one kernel pass: 28 B per element of HBM traffic."""

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3,
                 betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        if lr <= 0:
            msg = f"Invalid learning rate: {lr}"
            raise ValueError(msg)
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            msg = f"Invalid betas: {betas}"
            raise ValueError(msg)
        super().__init__(params, {"lr": lr, "betas": betas, "eps": eps})

    @torch.no_grad()
    def step(self, closure: Callable[[], torch.Tensor] | None = None):  # noqa: ANN201
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check_param(p)
                state = self.state[p]
                if not state:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p,
                                                           memory_format=torch.contiguous_format)
                state["step"] += 1
                t = state["step"]
                bias1 = 1.0 - beta1 ** t
                bias2 = 1.0 - beta2 ** t
                ops.adam_step(p, p.grad.contiguous(), state["exp_avg"], state["exp_avg_sq"],
                              beta1=beta1, beta2=beta2, eps=group["eps"],
                              step_size=group["lr"] / bias1, bias2_sqrt=bias2 ** 0.5)
        return loss


class FusedLBFGS(Optimizer):
    """L-BFGS with torch.optim.LBFGS semantics (``line_search_fn=None`` only).

    With ``max_iter == 1`` (the reference's setting) the whole update -- curvature-pair bookkeeping,
    two-loop recursion, step-length rule, tolerance tests, ``x += t*d`` -- is ONE C-ABI call that
    stays on the device (``stv_lbfgs_step``: three kernels, no host synchronisation), so a step costs
    one closure evaluation plus two passes over the stored history.  ``max_iter > 1`` keeps torch's
    host-driven control flow on the dot / axpy kernels.
    """

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1.0, max_iter: int = 20,  # noqa: PLR0913
                 max_eval: int | None = None, tolerance_grad: float = 1e-7,
                 tolerance_change: float = 1e-9, history_size: int = 100,
                 line_search_fn: str | None = None) -> None:
        if line_search_fn is not None:
            msg = "FusedLBFGS supports line_search_fn=None only (the reference's setting)"
            raise ValueError(msg)
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        super().__init__(params, {
            "lr": lr, "max_iter": max_iter, "max_eval": max_eval,
            "tolerance_grad": tolerance_grad, "tolerance_change": tolerance_change,
            "history_size": history_size, "line_search_fn": line_search_fn})
        self._image = _single_param(self)
        self._scratch: torch.Tensor | None = None
        self._scalars: torch.Tensor | None = None
        self._dev: dict[str, torch.Tensor] | None = None   # device-resident state (max_iter == 1)

    # -- small device helpers ----------------------------------------------------------------
    def _bufs(self) -> tuple[torch.Tensor, torch.Tensor]:
        if self._scratch is None:
            dev = self._image.device
            self._scratch = torch.empty(2 * nat.reduce_scratch_floats(), device=dev)
            self._scalars = torch.zeros(4, device=dev)
        return self._scratch, self._scalars

    def _dot(self, a: torch.Tensor, b: torch.Tensor) -> float:
        scratch, scalars = self._bufs()
        ops.dot(a, b, scratch, scalars[:1])
        return float(scalars[0].item())

    def _abs_stats(self, a: torch.Tensor) -> tuple[float, float]:
        scratch, scalars = self._bufs()
        ops.absmax_sum(a, scratch, scalars[:2])
        host = scalars[:2].tolist()
        return host[0], host[1]

    # -- device-resident single-iteration path ------------------------------------------------
    def _device_state(self) -> dict[str, torch.Tensor]:
        if self._dev is None:
            x = self._image
            n = x.numel()
            m = int(self.param_groups[0]["history_size"])
            stride = (n + 3) // 4 * 4
            dev = x.device
            self._dev = {
                "hist_s": torch.empty(m + 1, stride, device=dev, dtype=torch.float32),
                "hist_y": torch.empty(m + 1, stride, device=dev, dtype=torch.float32),
                "prev_g": torch.zeros(n, device=dev, dtype=torch.float32),
                "d": torch.zeros(n, device=dev, dtype=torch.float32),
                "work": torch.zeros(ops.lbfgs_workspace_floats(n, m), device=dev,
                                    dtype=torch.float32),
            }
        return self._dev

    def device_step(self, grad: torch.Tensor) -> None:
        """Apply one L-BFGS iteration for gradient ``grad`` entirely on the device."""
        group = self.param_groups[0]
        st = self._device_state()
        ops.lbfgs_step(self._image.detach().view(-1), grad.reshape(-1), st["hist_s"], st["hist_y"],
                       st["prev_g"], st["d"], st["work"], history=int(group["history_size"]),
                       lr=float(group["lr"]), tolerance_grad=float(group["tolerance_grad"]),
                       tolerance_change=float(group["tolerance_change"]))

    def device_counters(self) -> dict[str, int]:
        """n_iter / stored pairs / last-step flags (one small D2H read; for tests and logs)."""
        if self._dev is None:
            return {"n_iter": 0, "pairs": 0, "updated": 0, "stationary": 0}
        w = self._dev["work"][:8].view(torch.int32).tolist()
        return {"n_iter": w[0], "pairs": w[1], "updated": w[3], "stationary": w[4]}

    @torch.no_grad()
    def step(self, closure: Callable[[], torch.Tensor]):  # noqa: ANN201, C901, PLR0912, PLR0915
        group = self.param_groups[0]
        if group["max_iter"] == 1:
            x = self._image
            _check_param(x)
            with torch.enable_grad():
                orig_loss = closure()
            self.state[x]["func_evals"] = self.state[x].get("func_evals", 0) + 1
            if x.grad is None:
                msg = "closure did not produce a gradient for the image"
                raise RuntimeError(msg)
            self.device_step(x.grad.contiguous())
            return orig_loss
        lr, max_iter, max_eval = group["lr"], group["max_iter"], group["max_eval"]
        tol_grad, tol_change = group["tolerance_grad"], group["tolerance_change"]
        history_size = group["history_size"]
        x = self._image
        _check_param(x)
        closure = torch.enable_grad()(closure)

        state = self.state[x]
        state.setdefault("func_evals", 0)
        state.setdefault("n_iter", 0)

        orig_loss = closure()
        loss = float(orig_loss)
        evals = 1
        state["func_evals"] += 1
        if x.grad is None:
            msg = "closure did not produce a gradient for the image"
            raise RuntimeError(msg)
        g = x.grad.reshape(-1).contiguous()
        gmax, gl1 = self._abs_stats(g)
        if gmax <= tol_grad:  # already at a stationary point
            return orig_loss

        d: torch.Tensor | None = state.get("d")
        t: float = state.get("t", lr)
        ys_hist: list[torch.Tensor] = state.setdefault("old_dirs", [])
        ss_hist: list[torch.Tensor] = state.setdefault("old_stps", [])
        rho: list[float] = state.setdefault("ro", [])
        h_diag: float = state.get("H_diag", 1.0)
        prev_g: torch.Tensor | None = state.get("prev_flat_grad")
        prev_loss: float = state.get("prev_loss", loss)

        n_local = 0
        while n_local < max_iter:
            n_local += 1
            state["n_iter"] += 1

            # ---- search direction ----------------------------------------------------------
            if state["n_iter"] == 1:
                d = torch.empty_like(g)
                ops.scale(-1.0, g, d)
                ys_hist.clear()
                ss_hist.clear()
                rho.clear()
                h_diag = 1.0
            else:
                assert d is not None
                assert prev_g is not None
                y = g.clone()
                ops.axpy(-1.0, prev_g, y)           # y = g - g_prev
                s = torch.empty_like(d)
                ops.scale(t, d, s)                  # s = t * d
                ys = self._dot(y, s)
                if ys > 1e-10:
                    if len(ys_hist) == history_size:
                        ys_hist.pop(0)
                        ss_hist.pop(0)
                        rho.pop(0)
                    ys_hist.append(y)
                    ss_hist.append(s)
                    rho.append(1.0 / ys)
                    h_diag = ys / self._dot(y, y)
                # two-loop recursion
                m = len(ys_hist)
                alpha = [0.0] * m
                q = torch.empty_like(g)
                ops.scale(-1.0, g, q)
                for i in range(m - 1, -1, -1):
                    alpha[i] = self._dot(ss_hist[i], q) * rho[i]
                    ops.axpy(-alpha[i], ys_hist[i], q)
                ops.scale(h_diag, q, q)
                for i in range(m):
                    beta = self._dot(ys_hist[i], q) * rho[i]
                    ops.axpy(alpha[i] - beta, ss_hist[i], q)
                d = q

            if prev_g is None:
                prev_g = g.clone()
            else:
                prev_g.copy_(g)
            prev_loss = loss

            # ---- step length ---------------------------------------------------------------
            t = min(1.0, 1.0 / gl1) * lr if state["n_iter"] == 1 else lr

            gtd = self._dot(g, d)
            if gtd > -tol_change:  # not a descent direction (within tolerance)
                break

            ops.axpy(t, d, x.view(-1))              # x += t * d
            ls_evals = 0
            if n_local != max_iter:
                # re-evaluate at the new point (not reached with the reference's max_iter=1)
                loss = float(closure())
                g = x.grad.reshape(-1).contiguous()
                gmax, gl1 = self._abs_stats(g)
                ls_evals = 1
            evals += ls_evals
            state["func_evals"] += ls_evals

            # ---- termination ---------------------------------------------------------------
            if n_local == max_iter or evals >= max_eval:
                break
            if gmax <= tol_grad:
                break
            dmax, _ = self._abs_stats(d)
            if dmax * t <= tol_change:
                break
            if abs(loss - prev_loss) < tol_change:
                break

        state["d"] = d
        state["t"] = t
        state["H_diag"] = h_diag
        state["prev_flat_grad"] = prev_g
        state["prev_loss"] = prev_loss
        return orig_loss
