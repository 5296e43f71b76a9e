"""Execution engine: the truncated VGG19 feature stack, its losses and their input gradient as a
fixed program of sm_100a kernels over preallocated NHWC workspaces.

This is what runs underneath ``StyleContentModel.forward`` / ``loss.backward()`` (reference
core_model.py:297-328 and optimization.py:292,313).  The layer list is fused into *stages*
``conv [-> relu [-> pool]]``; ReLU is applied in the producing conv's epilogue (dual pre/post
outputs where a loss taps the pre-activation), the ReLU backward mask is applied in the epilogue
of the *next* layer's dgrad, and the loss gradients are accumulated in place into the tapped
layer's gradient buffer.  Any tap position the reference allows (conv, ReLU or pool outputs) is
supported; non-default positions take a few extra memory-bound passes.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field

import torch
from torch import nn

from . import _native as nat
from . import ops


# Default of VggLossEngine.compact_backward for engines created from now on (A/B measurements flip it
# before building a model).
DEFAULT_COMPACT_BACKWARD = True
DEFAULT_FUSE_STYLE_BWD = True
# conv1_1 forward on the tensor cores (TF32 operands, like cuDNN in the reference's CUDA path)
# instead of exact-fp32 CUDA cores.  Measured (profiles/r2_parity_report_2.log, r2_bench*_v8.json):
# +1.5 % steps/s at 512x512, +0.6 % at 1080p, but rounding the IMAGE to 10 mantissa bits raises the
# input-gradient deviation from the fp32 reference by ~20 % (1080p: 1.53e-2 -> 1.83e-2).  Parity
# comes first: off by default.
DEFAULT_FIRST_LAYER_TC = False


def _on_own_device(method):  # noqa: ANN001, ANN202
    """Run an engine entry point with the engine's GPU as the current CUDA device: the native
    launches use the current device's context (kernel attributes, SM count, stream), so a model on
    ``cuda:1`` must not launch while ``cuda:0`` is current (reference: ``--device cuda:1``)."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):  # noqa: ANN001, ANN002, ANN003, ANN202
        if torch.cuda.current_device() == self.device.index:
            return method(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapper


@dataclass
class _Stage:
    conv_idx: int
    cin: int
    cout: int
    weight: torch.Tensor            # torch layout [Cout, Cin, 3, 3]
    bias: torch.Tensor | None
    relu_idx: int | None = None     # index of the ReLU fused into this stage (None: out of range)
    pool_idx: int | None = None
    w_fwd: torch.Tensor | None = None
    w_dgrad: torch.Tensor | None = None


@dataclass
class _Workspace:
    """All device buffers of one image size; reused every step (CUDA-graph friendly)."""

    height: int
    width: int
    hw: list[tuple[int, int]] = field(default_factory=list)       # conv spatial size per stage
    pre: list[torch.Tensor | None] = field(default_factory=list)
    post: list[torch.Tensor | None] = field(default_factory=list)
    pool: list[torch.Tensor | None] = field(default_factory=list)
    d_y: list[torch.Tensor | None] = field(default_factory=list)
    d_post: list[torch.Tensor | None] = field(default_factory=list)
    d_pool: list[torch.Tensor | None] = field(default_factory=list)
    bits: list[torch.Tensor | None] = field(default_factory=list)   # ReLU sign bits (int32 words)
    code: list[torch.Tensor | None] = field(default_factory=list)   # pool argmax + gate nibbles
    gram_ws: list[torch.Tensor] = field(default_factory=list)
    s_mat: list[torch.Tensor] = field(default_factory=list)
    losses: torch.Tensor | None = None
    scratch: torch.Tensor | None = None
    grad_img: torch.Tensor | None = None
    generation: int = 0
    warm: tuple | None = None   # option signature of the last full forward + backward at this size (Engine.is_warm)
    nbytes: int = 0


class VggLossEngine:
    """Runs forward losses / input gradient for one model (weights + tap sets) on one device."""

    def __init__(self, layers: list[nn.Module], style_idx: list[int], content_idx: list[int],
                 device: torch.device) -> None:
        nat.require_device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.style_idx = sorted(set(style_idx))
        self.content_idx = sorted(set(content_idx))
        self._tapped = set(self.style_idx) | set(self.content_idx)
        with torch.cuda.device(device):
            self.stages = self._build_stages(layers)
        self._workspaces: dict[tuple[int, int], _Workspace] = {}
        self.style_targets: list[torch.Tensor] | None = None          # [C, C] each
        self.content_targets_nhwc: list[torch.Tensor] | None = None   # [h, w, C] each
        # Loss kernels (Gram, content MSE, their backward seeds) are memory-bound and independent of
        # the next conv layers: they run on a side stream, forked/joined with events, so they
        # overlap the tensor-core work (the fork/join structure is preserved by CUDA-graph capture).
        self._side = torch.cuda.Stream(device=device)
        self.overlap_losses = True
        # Backward bookkeeping as bits instead of fp32 re-reads: the forward epilogues record the ReLU
        # sign bits (1 bit / element) and the pool argmax + gate (4 bits / pooled element); the dgrad
        # epilogues gate from the bits and route pooled gradients themselves (no pool-backward
        # kernel, no full-resolution activation kept for pooled layers).  False = fp32 re-reads, kept
        # for A/B measurements; stages whose ReLU / pool output is tapped by a loss always use those.
        self.compact_backward = DEFAULT_COMPACT_BACKWARD
        self.fuse_style_bwd = DEFAULT_FUSE_STYLE_BWD
        self.first_layer_tc = DEFAULT_FIRST_LAYER_TC

    # ------------------------------------------------------------------ program construction
    def _build_stages(self, layers: list[nn.Module]) -> list[_Stage]:
        stages: list[_Stage] = []
        i = 0
        n = len(layers)
        while i < n:
            layer = layers[i]
            if not isinstance(layer, nn.Conv2d):
                msg = f"layer {i}: expected Conv2d to start a stage, found {type(layer).__name__}"
                raise TypeError(msg)
            if layer.kernel_size != (3, 3) or layer.padding != (1, 1) or layer.stride != (1, 1):
                msg = f"layer {i}: only 3x3 / stride 1 / pad 1 convolutions are supported"
                raise ValueError(msg)
            w = layer.weight.detach().to(self.device, torch.float32).contiguous()
            b = None if layer.bias is None else \
                layer.bias.detach().to(self.device, torch.float32).contiguous()
            st = _Stage(conv_idx=i, cin=w.shape[1], cout=w.shape[0], weight=w, bias=b)
            i += 1
            if i < n and isinstance(layers[i], nn.ReLU):
                st.relu_idx = i
                i += 1
                if i < n and isinstance(layers[i], nn.MaxPool2d):
                    pool = layers[i]
                    ks = pool.kernel_size if isinstance(pool.kernel_size, tuple) \
                        else (pool.kernel_size,) * 2
                    sd = pool.stride if isinstance(pool.stride, tuple) else (pool.stride,) * 2
                    if ks != (2, 2) or sd != (2, 2) or pool.ceil_mode:
                        msg = f"layer {i}: only MaxPool2d(2, 2) floor-mode is supported"
                        raise ValueError(msg)
                    st.pool_idx = i
                    i += 1
            elif i < n:
                msg = f"layer {i}: a convolution must be followed by ReLU in this network"
                raise TypeError(msg)
            stages.append(st)
        if not stages:
            msg = "no layers to run: style/content layer indices select nothing"
            raise ValueError(msg)
        first = stages[0]
        if first.cin != 3 or first.cout != 64:
            msg = f"first convolution must be 3->64 channels (got {first.cin}->{first.cout})"
            raise ValueError(msg)
        first.w_dgrad = ops.pack_first_dgrad_rows(first.weight)  # [3, 16, 64]: x taps folded into N
        for st in stages[1:]:
            if st.cin % 64 or st.cout % 64:
                msg = (f"layer {st.conv_idx}: tensor-core conv needs channel counts that are "
                       f"multiples of 64 (got {st.cin}->{st.cout})")
                raise ValueError(msg)
            st.w_fwd, st.w_dgrad = ops.pack_conv_weights(st.weight)
        return stages

    def _out_hw(self, height: int, width: int) -> list[tuple[int, int]]:
        sizes = []
        h, w = height, width
        for st in self.stages:
            sizes.append((h, w))
            if st.pool_idx is not None:
                h, w = h // 2, w // 2
                if h < 1 or w < 1:
                    msg = f"image {height}x{width} is too small for the pooling depth"
                    raise ValueError(msg)
        return sizes

    def _workspace(self, height: int, width: int, *, with_grad: bool) -> _Workspace:
        key = (height, width)
        ws = self._workspaces.get(key)
        if ws is None:
            ws = _Workspace(height=height, width=width)
            ws.hw = self._out_hw(height, width)
            dev = self.device

            def buf(*shape: int) -> torch.Tensor:
                t = torch.empty(*shape, device=dev, dtype=torch.float32)
                ws.nbytes += t.numel() * 4
                return t

            for s, (st, (h, w)) in enumerate(zip(self.stages, ws.hw)):
                need_pre = st.conv_idx in self._tapped or st.relu_idx is None
                ws.pre.append(buf(h, w, st.cout) if need_pre else None)
                unpool = self._uses_unpool(s)
                ws.post.append(buf(h, w, st.cout) if st.relu_idx is not None and not unpool
                               else None)
                ws.pool.append(buf(h // 2, w // 2, st.cout) if st.pool_idx is not None else None)
                ws.bits.append(ops.relu_bits_buffer(h, w, st.cout, dev)
                               if self._uses_bits(s) else None)
                ws.code.append(ops.pool_code_buffer(h, w, st.cout, dev) if unpool else None)
                ws.d_y.append(None)
                ws.d_post.append(None)
                ws.d_pool.append(None)
            for idx in self.style_idx:
                t = self._tap_tensor(ws, idx)
                c = t.shape[-1]
                ws.gram_ws.append(ops.gram_workspace(t.shape[0] * t.shape[1], c, dev))
                ws.s_mat.append(buf(c, c))
            ws.losses = torch.zeros(len(self.style_idx) + len(self.content_idx), device=dev,
                                    dtype=torch.float32)
            ws.scratch = buf(2 * nat.reduce_scratch_floats())
            self._workspaces[key] = ws
        if with_grad and ws.grad_img is None:
            dev = self.device
            for s, (st, (h, w)) in enumerate(zip(self.stages, ws.hw)):
                # zero-filled: the un-pooling dgrad never writes a last row / column that floor-mode
                # pooling dropped (its gradient is zero)
                ws.d_y[s] = torch.zeros(h, w, st.cout, device=dev, dtype=torch.float32)
                ws.nbytes += ws.d_y[s].numel() * 4
                if st.pool_idx is not None and not self._uses_unpool(s):
                    ws.d_pool[s] = torch.empty(h // 2, w // 2, st.cout, device=dev,
                                               dtype=torch.float32)
                if st.relu_idx is not None and st.relu_idx in self._tapped:
                    ws.d_post[s] = torch.empty(h, w, st.cout, device=dev, dtype=torch.float32)
            ws.grad_img = torch.empty(1, 3, height, width, device=dev, dtype=torch.float32)
        return ws

    def _uses_bits(self, s: int) -> bool:
        """Stage ``s`` = conv -> ReLU (no pool, ReLU output not tapped) feeding another conv: its ReLU
        backward is applied from recorded sign bits in the next layer's dgrad epilogue."""
        st = self.stages[s]
        return (self.compact_backward and st.relu_idx is not None and st.pool_idx is None
                and st.relu_idx not in self._tapped and s + 1 < len(self.stages))

    def _split_relu(self, s: int, ws: _Workspace) -> bool:
        st = self.stages[s]
        return (self.compact_backward and s > 0 and ws.pre[s] is not None and ws.post[s] is not None
                and st.pool_idx is None and st.cout >= 256 and st.cout % 32 == 0)

    def _fuses_style_bwd(self, s: int) -> bool:
        """Stage ``s``: conv output tapped by a style loss only, gated by sign bits, 64 or 128
        channels wide -- its Gram backward runs as a second accumulator of the next layer's dgrad."""
        st = self.stages[s]
        return (self._uses_bits(s) and st.conv_idx in self.style_idx
                and st.conv_idx not in self.content_idx and st.cout in (64, 128)
                and self.fuse_style_bwd)

    def _uses_unpool(self, s: int) -> bool:
        """Stage ``s`` = conv -> ReLU -> pool with no loss tap, feeding another conv: the pool + ReLU
        backward is applied from recorded codes in the next layer's dgrad epilogue."""
        st = self.stages[s]
        return (self.compact_backward and s > 0 and st.pool_idx is not None
                and not self._stage_taps(st) and s + 1 < len(self.stages))

    def is_warm(self, height: int, width: int) -> bool:
        """True once a full forward + backward has run at this image size: workspaces are allocated
        and the kernels' one-time attribute calls are done, so a step can be captured into a CUDA
        graph without a dry run."""
        ws = self._workspaces.get((height, width))
        return ws is not None and ws.warm == self._options_signature()

    def _options_signature(self) -> tuple:
        return (self.compact_backward, self.fuse_style_bwd, self.first_layer_tc, self.overlap_losses)

    def grad_buffer(self, height: int, width: int) -> torch.Tensor:
        """The NCHW buffer ``backward_losses`` writes for this image size."""
        return self._workspace(height, width, with_grad=True).grad_img

    def generation_of(self, height: int, width: int) -> int:
        ws = self._workspaces.get((height, width))
        return ws.generation if ws else 0

    def release_workspaces(self) -> None:
        self._workspaces.clear()

    def workspace_bytes(self, height: int, width: int) -> int:
        ws = self._workspaces.get((height, width))
        return ws.nbytes if ws else 0

    def _stage_of(self, idx: int) -> tuple[int, str]:
        for s, st in enumerate(self.stages):
            if idx == st.conv_idx:
                return s, "pre"
            if idx == st.relu_idx:
                return s, "post"
            if idx == st.pool_idx:
                return s, "pool"
        msg = f"layer index {idx} is not part of the truncated network"
        raise KeyError(msg)

    def _tap_tensor(self, ws: _Workspace, idx: int) -> torch.Tensor:
        s, kind = self._stage_of(idx)
        t = getattr(ws, kind)[s]
        assert t is not None
        return t

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _check_image(x: torch.Tensor) -> tuple[int, int]:
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 3:
            msg = (f"expected an image tensor of shape [1, 3, H, W], got {tuple(x.shape)} "
                   "(independent images are independent jobs; the reference folds batch into "
                   "the Gram channels)")
            raise ValueError(msg)
        if x.dtype != torch.float32 or not x.is_cuda:
            msg = f"expected a float32 CUDA tensor, got {x.dtype} on {x.device}"
            raise ValueError(msg)
        return int(x.shape[2]), int(x.shape[3])

    def _run_stack(self, x: torch.Tensor, ws: _Workspace, after_stage=None) -> None:  # noqa: ANN001
        cur: torch.Tensor | None = None
        for s, st in enumerate(self.stages):
            if s == 0 and self.first_layer_tc:
                ops.conv3x3_first_fwd_tc(x, st.weight, st.bias, ws.pre[0], ws.post[0],
                                         round_pre=self._round_pre(st), out_bits=ws.bits[0])
            elif s == 0:
                ops.conv3x3_first_fwd(x, st.weight, st.bias, ws.pre[0], ws.post[0],
                                      round_pre=self._round_pre(st), out_bits=ws.bits[0])
            elif ws.code[s] is not None:
                # conv + ReLU + MaxPool2d in one launch; only the pooled map and the 4-bit pool
                # codes reach memory (the backward pass needs nothing else of this layer)
                ops.conv3x3_fwd(cur, st.w_fwd, st.bias, ws.pre[s], None,
                                round_pre=self._round_pre(st), out_pool=ws.pool[s],
                                out_code=ws.code[s])
            elif self._split_relu(s, ws):
                # 256-wide tiles cannot coalesce two output tensors (no shared memory left for the
                # staging tiles) and a dual-output epilogue of scattered stores costs more than the
                # whole MMA work of the layer: the conv stores the tapped pre-activation only, a
                # memory-bound pass makes the ReLU'd operand + sign bits from it
                ops.conv3x3_fwd(cur, st.w_fwd, st.bias, ws.pre[s], None,
                                round_pre=self._round_pre(st))
                ops.relu_fwd_bits(ws.pre[s], ws.post[s], ws.bits[s])
            else:
                # the MaxPool2d after conv1_2 / 2_2 / 3_4 / 4_4 is computed in the conv epilogue
                fused_pool = ws.pool[s] if st.pool_idx is not None and ws.post[s] is not None else None
                ops.conv3x3_fwd(cur, st.w_fwd, st.bias, ws.pre[s], ws.post[s],
                                round_pre=self._round_pre(st), out_pool=fused_pool,
                                out_bits=ws.bits[s] if fused_pool is None else None)
            cur = ws.post[s] if ws.post[s] is not None else ws.pre[s]
            if st.pool_idx is not None:
                if ws.code[s] is None and (s == 0 or ws.post[s] is None):
                    ops.maxpool2_fwd(ws.post[s], ws.pool[s])
                cur = ws.pool[s]
            if after_stage is not None:
                after_stage(s, st)

    def _round_pre(self, st: _Stage) -> bool:
        """A conv output tapped by style losses only is consumed solely by tensor-core kernels
        (Gram, Gram backward): store it tf32-rounded so their operand truncation is exact.  Content
        taps keep the exact fp32 value (the content MSE is fp32 arithmetic)."""
        return st.conv_idx in self.style_idx and st.conv_idx not in self.content_idx

    def _stage_taps(self, st: _Stage) -> list[int]:
        return [i for i in (st.conv_idx, st.relu_idx, st.pool_idx)
                if i is not None and i in self._tapped]

    @_on_own_device
    def compute_targets(self, style_img: torch.Tensor, content_img: torch.Tensor) -> None:
        """reference core_model.py:192-232 -- Grams of the style image, features of the content
        image, produced by the same kernels as the per-step forward."""
        sh, sw = self._check_image(style_img)
        ch, cw = self._check_image(content_img)
        ws = self._workspace(sh, sw, with_grad=False)
        self._run_stack(style_img.contiguous(), ws)
        # Existing target tensors of matching shape are overwritten IN PLACE, so a captured step
        # graph (which holds their addresses) stays valid when the next job's targets are loaded.
        old_grams = self.style_targets or []
        grams = []
        for k, idx in enumerate(self.style_idx):
            t = self._tap_tensor(ws, idx)
            c = t.shape[-1]
            g = old_grams[k] if k < len(old_grams) and old_grams[k].shape == (c, c) else \
                torch.empty(c, c, device=self.device, dtype=torch.float32)
            ops.gram_loss_fwd(t, ws.gram_ws[k], gram_out=g)
            grams.append(g)
        if (sh, sw) != (ch, cw):
            del self._workspaces[(sh, sw)]  # style-sized buffers are never needed again
        ws = self._workspace(ch, cw, with_grad=False)
        self._run_stack(content_img.contiguous(), ws)
        old_feats = self.content_targets_nhwc or []
        feats = []
        for k, idx in enumerate(self.content_idx):
            t = self._tap_tensor(ws, idx)
            if k < len(old_feats) and old_feats[k].shape == t.shape:
                old_feats[k].copy_(t)
                feats.append(old_feats[k])
            else:
                feats.append(t.clone())
        self.style_targets = grams
        self.content_targets_nhwc = feats

    @_on_own_device
    def forward_losses(self, x: torch.Tensor) -> tuple[torch.Tensor, int]:
        """Run the stack and all losses; returns (view of the loss buffer, workspace generation).
        Loss order: style layers ascending, then content layers ascending."""
        if self.style_targets is None or self.content_targets_nhwc is None:
            msg = "targets must be set before computing losses."
            raise RuntimeError(msg)
        h, w = self._check_image(x)
        ws = self._workspace(h, w, with_grad=False)
        ns = len(self.style_idx)
        main = torch.cuda.current_stream(self.device)
        side = self._side if self.overlap_losses else main

        def losses_of(idx: int) -> None:
            t = self._tap_tensor(ws, idx)
            if idx in self.style_idx:
                k = self.style_idx.index(idx)
                ops.gram_loss_fwd(t, ws.gram_ws[k], target=self.style_targets[k],
                                  s_out=ws.s_mat[k], loss_out=ws.losses[k:k + 1])
            if idx in self.content_idx:
                k = self.content_idx.index(idx)
                tgt = self.content_targets_nhwc[k]
                if tgt.shape != t.shape:
                    msg = (f"content target shape {tuple(tgt.shape)} does not match features "
                           f"{tuple(t.shape)} at layer {idx}")
                    raise RuntimeError(msg)
                ops.content_loss_fwd(t, tgt, ws.scratch, ws.losses[ns + k:ns + k + 1])

        def after_stage(_s: int, st: _Stage) -> None:
            taps = self._stage_taps(st)
            if not taps:
                return
            if side is not main:
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
            with torch.cuda.stream(side):
                for idx in taps:
                    losses_of(idx)

        self._run_stack(x if x.is_contiguous() else x.contiguous(), ws, after_stage)
        if side is not main:
            main.wait_stream(side)
        ws.generation += 1
        return ws.losses, ws.generation

    # ------------------------------------------------------------------ backward
    def _tap_grads(self, ws: _Workspace, idx: int | None, tensor: torch.Tensor | None,
                   out: torch.Tensor, grad_w: torch.Tensor, accumulate: bool) -> bool:
        """Accumulate d(loss)/d(tensor) of every loss tapping layer ``idx`` into ``out``."""
        if idx is None or idx not in self._tapped:
            return accumulate
        assert tensor is not None
        ns = len(self.style_idx)
        if idx in self.style_idx:
            k = self.style_idx.index(idx)
            ops.style_bwd(tensor, ws.s_mat[k], grad_w[k:k + 1], out, accumulate=accumulate)
            accumulate = True
        if idx in self.content_idx:
            k = self.content_idx.index(idx)
            ops.content_loss_bwd(tensor, self.content_targets_nhwc[k], grad_w[ns + k:ns + k + 1],
                                 out, accumulate=accumulate)
            accumulate = True
        return accumulate

    @_on_own_device
    def backward_losses(self, height: int, width: int, grad_w: torch.Tensor,
                        generation: int | None = None) -> torch.Tensor:
        """Input gradient  sum_k grad_w[k] * d(loss_k)/dx  using the activations of the most
        recent ``forward_losses`` at this image size.  Returns the workspace's NCHW buffer."""
        ws = self._workspace(height, width, with_grad=True)
        if generation is not None and generation != ws.generation:
            msg = ("backward called with stale activations: another forward pass ran on this "
                   "model (same image size) since the forward being differentiated")
            raise RuntimeError(msg)
        grad_w = grad_w.to(torch.float32).contiguous()
        n = len(self.stages)
        # Loss gradients at pre-activation taps depend only on the forward pass: issue them all up
        # front on the side stream; the dgrad that accumulates onto one waits for its event.
        main = torch.cuda.current_stream(self.device)
        early: dict[int, torch.cuda.Event] = {}
        if self.overlap_losses:
            fork = torch.cuda.Event()
            fork.record(main)
            self._side.wait_event(fork)
            with torch.cuda.stream(self._side):
                for s in range(n - 1, -1, -1):
                    st = self.stages[s]
                    tap_post = st.relu_idx is not None and st.relu_idx in self._tapped
                    if st.pool_idx is None and not tap_post and st.conv_idx in self._tapped \
                            and not self._fuses_style_bwd(s):
                        self._tap_grads(ws, st.conv_idx, ws.pre[s], ws.d_y[s], grad_w, False)
                        done = torch.cuda.Event()
                        done.record(self._side)
                        early[s] = done
        for s in range(n - 1, -1, -1):
            st = self.stages[s]
            down = self.stages[s + 1] if s + 1 < n else None
            d_y = ws.d_y[s]
            tap_post = st.relu_idx is not None and st.relu_idx in self._tapped
            if ws.code[s] is not None:
                # dgrad of the next conv at pooled resolution, pool + ReLU backward in its epilogue
                ops.conv3x3_dgrad_unpool(ws.d_y[s + 1], down.w_dgrad, ws.code[s], d_y)
            elif st.pool_idx is not None:
                d_pool = ws.d_pool[s]
                acc = self._tap_grads(ws, st.pool_idx, ws.pool[s], d_pool, grad_w, False)
                if down is not None:
                    ops.conv3x3_dgrad(ws.d_y[s + 1], down.w_dgrad, d_pool, accumulate=acc)
                if tap_post:
                    d_post = ws.d_post[s]
                    ops.maxpool2_bwd(d_pool, ws.post[s], d_post, relu_mask=False)
                    self._tap_grads(ws, st.relu_idx, ws.post[s], d_post, grad_w, True)
                    acc2 = self._tap_grads(ws, st.conv_idx, ws.pre[s], d_y, grad_w, False)
                    ops.relu_bwd(d_post, ws.post[s], d_y, accumulate=acc2)
                else:
                    ops.maxpool2_bwd(d_pool, ws.post[s], d_y, relu_mask=True)
                    self._tap_grads(ws, st.conv_idx, ws.pre[s], d_y, grad_w, True)
            elif st.relu_idx is not None:
                if tap_post:
                    d_post = ws.d_post[s]
                    acc = self._tap_grads(ws, st.relu_idx, ws.post[s], d_post, grad_w, False)
                    if down is not None:
                        ops.conv3x3_dgrad(ws.d_y[s + 1], down.w_dgrad, d_post, accumulate=acc)
                    acc2 = self._tap_grads(ws, st.conv_idx, ws.pre[s], d_y, grad_w, False)
                    ops.relu_bwd(d_post, ws.post[s], d_y, accumulate=acc2)
                elif self._fuses_style_bwd(s):
                    # ReLU gate, dgrad and this layer's Gram backward in ONE launch
                    k = self.style_idx.index(st.conv_idx)
                    ops.conv3x3_dgrad_style(ws.d_y[s + 1], down.w_dgrad, d_y, relu_bits=ws.bits[s],
                                            feat=ws.pre[s], s_mat=ws.s_mat[k],
                                            grad_w=grad_w[k:k + 1])
                else:
                    if s in early:
                        main.wait_event(early[s])
                        acc = True
                    else:
                        acc = self._tap_grads(ws, st.conv_idx, ws.pre[s], d_y, grad_w, False)
                    # ReLU backward mask fused into the dgrad epilogue
                    if ws.bits[s] is not None:
                        ops.conv3x3_dgrad(ws.d_y[s + 1], down.w_dgrad, d_y, relu_bits=ws.bits[s],
                                          accumulate=acc)
                    else:
                        ops.conv3x3_dgrad(ws.d_y[s + 1], down.w_dgrad, d_y, relu_src=ws.post[s],
                                          accumulate=acc)
            elif s in early:
                main.wait_event(early[s])
            else:
                self._tap_grads(ws, st.conv_idx, ws.pre[s], d_y, grad_w, False)
        ops.conv3x3_first_dgrad_rows(ws.d_y[0], self.stages[0].w_dgrad, ws.grad_img)
        # every buffer of this size exists and every kernel variant these options select has been launched
        ws.warm = self._options_signature()
        return ws.grad_img

    # ------------------------------------------------------------------ introspection
    @_on_own_device
    def tap_features_nchw(self, x: torch.Tensor) -> list[torch.Tensor]:
        """Tapped activations (ascending layer index) as NCHW copies -- for tests."""
        h, w = self._check_image(x)
        ws = self._workspace(h, w, with_grad=False)
        self._run_stack(x.contiguous(), ws)
        ws.generation += 1
        return [ops.nhwc_to_nchw(self._tap_tensor(ws, idx)) for idx in sorted(self._tapped)]

    def flops_per_step(self, height: int, width: int) -> dict[str, float]:
        """Algorithmic FLOPs of one forward + backward at this size."""
        sizes = self._out_hw(height, width)
        conv = sum(2.0 * 9 * st.cin * st.cout * h * w for st, (h, w) in zip(self.stages, sizes))
        gram_tri = 0.0
        gram_full = 0.0
        for idx in self.style_idx:
            s, kind = self._stage_of(idx)
            h, w = sizes[s]
            if kind == "pool":
                h, w = h // 2, w // 2
            c = self.stages[s].cout
            gram_full += 2.0 * c * c * h * w
            gram_tri += 1.0 * c * (c + 1) * h * w
        return {"conv_fwd": conv, "conv_dgrad": conv, "gram_fwd_full": gram_full,
                "gram_fwd_tri": gram_tri, "style_bwd": gram_full,
                "total_tri": 2 * conv + gram_tri + gram_full}
