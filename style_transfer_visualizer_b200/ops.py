"""Tensor-level wrappers over the C-ABI kernels (one Python function per entry point).

All functions enqueue on torch's current CUDA stream and never synchronise, so sequences of them
can be captured into a CUDA graph.  Activations are NHWC fp32 tensors of shape ``[H, W, C]``; the
image and its gradient are NCHW ``[1, 3, H, W]`` as in the reference.
"""
from __future__ import annotations

import torch

from . import _native as nat

GRAM_MATRIX_CLAMP_MAX = 5e5  # reference constants.py:15


def _chk(t: torch.Tensor, name: str) -> None:
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        msg = f"{name}: expected a contiguous fp32 CUDA tensor, got {t.dtype} on {t.device}"
        raise ValueError(msg)


def _s(t: torch.Tensor) -> int:
    """Stream argument of a launch on ``t``'s device (and makes ``nat.call`` run it there)."""
    return nat.stream_for_call(t.device)


def pack_conv_weights(w: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """torch ``[Cout, Cin, 3, 3]`` -> (fwd ``[9, Cout, Cin]``, dgrad ``[9, Cin, Cout]`` flipped)."""
    _chk(w, "w")
    cout, cin = w.shape[0], w.shape[1]
    wf = torch.empty(9, cout, cin, device=w.device, dtype=torch.float32)
    wd = torch.empty(9, cin, cout, device=w.device, dtype=torch.float32)
    nat.call("stv_pack_conv_weights", nat.ptr(w), nat.ptr(wf), nat.ptr(wd), cout, cin, _s(w))
    return wf, wd


def relu_bits_buffer(h: int, w: int, channels: int, device: torch.device) -> torch.Tensor:
    """``[H, W, C/32]`` int32 words: sign bits of a post-ReLU activation (``stv_conv3x3_fwd_bits``)."""
    return torch.empty(h, w, channels // 32, device=device, dtype=torch.int32)


def pool_code_buffer(h: int, w: int, channels: int, device: torch.device) -> torch.Tensor:
    """``[H, W, C/32]`` int32 words: pool + ReLU backward routing bits (``stv_conv3x3_fwd_pool_code``)."""
    return torch.empty(h, w, channels // 32, device=device, dtype=torch.int32)


def conv3x3_first_fwd(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None,
                      out_pre: torch.Tensor | None, out_post: torch.Tensor | None, *,
                      round_pre: bool = False, out_bits: torch.Tensor | None = None) -> None:
    _chk(img, "img")
    h, wd = img.shape[-2], img.shape[-1]
    if out_bits is not None:
        nat.call("stv_conv3x3_first_fwd_bits", nat.ptr(img), nat.ptr(w), nat.ptr(bias), h, wd,
                 w.shape[0], nat.ptr(out_pre), nat.ptr(out_post), nat.ptr(out_bits),
                 int(round_pre), _s(img))
        return
    nat.call("stv_conv3x3_first_fwd", nat.ptr(img), nat.ptr(w), nat.ptr(bias), h, wd, w.shape[0],
             nat.ptr(out_pre), nat.ptr(out_post), int(round_pre), _s(img))


def conv3x3_first_fwd_tc(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None,  # noqa: PLR0913
                         out_pre: torch.Tensor | None, out_post: torch.Tensor | None, *,
                         rows: int | None = None, in_row0: int = 0, round_pre: bool = False,
                         out_bits: torch.Tensor | None = None) -> None:
    """conv1_1 forward on the tensor cores (``stv_conv3x3_first_fwd_tc``); ``img`` may be a haloed
    band ``[1, 3, in_rows, W]`` with ``rows`` output rows starting at input row ``in_row0``."""
    _chk(img, "img")
    in_rows = int(img.shape[-2])
    nat.call("stv_conv3x3_first_fwd_tc", nat.ptr(img), nat.ptr(w), nat.ptr(bias),
             in_rows if rows is None else rows, int(img.shape[-1]), int(w.shape[0]), in_rows,
             in_row0, nat.ptr(out_pre), nat.ptr(out_post), nat.ptr(out_bits), int(round_pre),
             _s(img))


def conv3x3_fwd(x: torch.Tensor, w_fwd: torch.Tensor, bias: torch.Tensor | None,  # noqa: PLR0913
                out_pre: torch.Tensor | None, out_post: torch.Tensor | None, *,
                round_pre: bool = False, out_pool: torch.Tensor | None = None,
                out_bits: torch.Tensor | None = None,
                out_code: torch.Tensor | None = None) -> None:
    """3x3 conv + bias (+ReLU into out_post); with ``out_pool`` also the 2x2 max pool of out_post,
    computed in the conv epilogue (``stv_conv3x3_fwd_pool``).  ``out_bits`` / ``out_code`` record the
    ReLU sign bits / the pool argmax codes for the backward pass."""
    _chk(x, "x")
    h, wd, cin = x.shape
    cout = w_fwd.shape[1]
    if out_code is not None:
        nat.call("stv_conv3x3_fwd_pool_code", nat.ptr(x), nat.ptr(w_fwd), nat.ptr(bias), h, wd, cin,
                 cout, nat.ptr(out_pre), nat.ptr(out_post), nat.ptr(out_pool), nat.ptr(out_code),
                 int(round_pre), _s(x))
        return
    if out_pool is not None:
        nat.call("stv_conv3x3_fwd_pool", nat.ptr(x), nat.ptr(w_fwd), nat.ptr(bias), h, wd, cin, cout,
                 nat.ptr(out_pre), nat.ptr(out_post), nat.ptr(out_pool), int(round_pre), _s(x))
        return
    if out_bits is not None:
        nat.call("stv_conv3x3_fwd_bits", nat.ptr(x), nat.ptr(w_fwd), nat.ptr(bias), h, wd, cin, cout,
                 nat.ptr(out_pre), nat.ptr(out_post), nat.ptr(out_bits), int(round_pre), _s(x))
        return
    nat.call("stv_conv3x3_fwd", nat.ptr(x), nat.ptr(w_fwd), nat.ptr(bias), h, wd, cin, cout,
             nat.ptr(out_pre), nat.ptr(out_post), int(round_pre), _s(x))


def conv3x3_dgrad(dy: torch.Tensor, w_dgrad: torch.Tensor, dx: torch.Tensor,
                  relu_src: torch.Tensor | None = None, *, accumulate: bool = False,
                  relu_bits: torch.Tensor | None = None) -> None:
    """Input gradient of a 3x3 conv; the ReLU gate comes from ``relu_src`` (fp32 activation) or from
    ``relu_bits`` (its recorded sign bits)."""
    _chk(dy, "dy")
    h, wd, cout = dy.shape
    cin = w_dgrad.shape[1]
    if relu_bits is not None:
        nat.call("stv_conv3x3_dgrad_bits", nat.ptr(dy), nat.ptr(w_dgrad), h, wd, cout, cin,
                 nat.ptr(relu_bits), int(accumulate), nat.ptr(dx), _s(dy))
        return
    nat.call("stv_conv3x3_dgrad", nat.ptr(dy), nat.ptr(w_dgrad), h, wd, cout, cin,
             nat.ptr(relu_src), int(accumulate), nat.ptr(dx), _s(dy))


def conv3x3_dgrad_style(dy: torch.Tensor, w_dgrad: torch.Tensor, dx: torch.Tensor, *,  # noqa: PLR0913
                        relu_bits: torch.Tensor | None, feat: torch.Tensor, s_mat: torch.Tensor,
                        grad_w: torch.Tensor) -> None:
    """``dx = bits .* dgrad(dy) + grad_w * feat @ s_mat``: conv input gradient and the Gram backward of
    the style-tapped layer in one launch (``stv_conv3x3_dgrad_bits_style``)."""
    _chk(dy, "dy")
    h, wd, cout = dy.shape
    cin = w_dgrad.shape[1]
    nat.call("stv_conv3x3_dgrad_bits_style", nat.ptr(dy), nat.ptr(w_dgrad), h, wd, cout, cin,
             nat.ptr(relu_bits), nat.ptr(feat), nat.ptr(s_mat), nat.ptr(grad_w), nat.ptr(dx), _s(dy))


def conv3x3_dgrad_unpool(dy: torch.Tensor, w_dgrad: torch.Tensor, pool_code: torch.Tensor,
                         dx: torch.Tensor) -> None:
    """dgrad at pooled resolution fused with the 2x2 max-pool + ReLU backward: ``dx`` ``[H2, W2, Cin]``
    is the gradient at the resolution before the pool (``stv_conv3x3_dgrad_unpool``)."""
    _chk(dy, "dy")
    h, wd, cout = dy.shape
    cin = w_dgrad.shape[1]
    h2, w2 = dx.shape[0], dx.shape[1]
    nat.call("stv_conv3x3_dgrad_unpool", nat.ptr(dy), nat.ptr(w_dgrad), h, wd, cout, cin,
             nat.ptr(pool_code), h2, w2, nat.ptr(dx), _s(dy))


def conv3x3_desc(x: torch.Tensor, w_packed: torch.Tensor, *, rows: int, x_row0: int = 0,  # noqa: PLR0913
                 taps: int = 9, bias: torch.Tensor | None = None,
                 add_src: torch.Tensor | None = None, out_pre: torch.Tensor | None = None,
                 out_post: torch.Tensor | None = None, round_flags: int = 0,
                 out_pool: torch.Tensor | None = None, out_bits: torch.Tensor | None = None,
                 out_code: torch.Tensor | None = None, mask_bits: torch.Tensor | None = None,
                 unpool_code: torch.Tensor | None = None, unpool_hw: tuple[int, int] = (0, 0),
                 style_x: torch.Tensor | None = None, style_s: torch.Tensor | None = None,
                 style_alpha: torch.Tensor | None = None) -> bool:
    """General form of the tensor-core conv (``stv_conv3x3_desc``): ``x`` may be a haloed buffer
    ``[x_rows, W, C]`` whose row ``x_row0`` lines up with output row 0; the outputs cover ``rows``
    rows.  Returns False (nothing launched) when a requested fused style backward is not available
    for the shape."""
    _chk(x, "x")
    d = nat.ConvDesc()
    d.x, d.w_packed = nat.ptr(x), nat.ptr(w_packed)
    d.H, d.W, d.C, d.N, d.taps = rows, int(x.shape[1]), int(x.shape[2]), int(w_packed.shape[1]), taps
    d.x_rows, d.x_row0 = int(x.shape[0]), x_row0
    d.bias, d.add_src = nat.ptr(bias), nat.ptr(add_src)
    d.out_pre, d.out_post, d.round_flags = nat.ptr(out_pre), nat.ptr(out_post), round_flags
    d.out_pool, d.out_bits, d.out_code = nat.ptr(out_pool), nat.ptr(out_bits), nat.ptr(out_code)
    d.mask_bits, d.unpool_code = nat.ptr(mask_bits), nat.ptr(unpool_code)
    d.H2, d.W2 = unpool_hw
    d.style_x, d.style_s, d.style_alpha = nat.ptr(style_x), nat.ptr(style_s), nat.ptr(style_alpha)
    import ctypes

    return nat.call_status("stv_conv3x3_desc", ctypes.byref(d), _s(x)) == 0


def conv3x3_first_fwd_band(img: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None,  # noqa: PLR0913
                           out_pre: torch.Tensor | None, out_post: torch.Tensor | None, *,
                           rows: int, in_row0: int, round_pre: bool = False,
                           out_bits: torch.Tensor | None = None) -> None:
    """conv1_1 forward of a haloed NCHW image band ``[1, 3, in_rows, W]`` (``rows`` output rows)."""
    _chk(img, "img")
    nat.call("stv_conv3x3_first_fwd_band", nat.ptr(img), nat.ptr(w), nat.ptr(bias), rows,
             int(img.shape[-1]), int(w.shape[0]), int(img.shape[-2]), in_row0, nat.ptr(out_pre),
             nat.ptr(out_post), nat.ptr(out_bits), int(round_pre), _s(img))


def conv3x3_first_dgrad(dy: torch.Tensor, w: torch.Tensor, dimg: torch.Tensor) -> None:
    _chk(dy, "dy")
    h, wd, cout = dy.shape
    nat.call("stv_conv3x3_first_dgrad", nat.ptr(dy), nat.ptr(w), h, wd, cout, nat.ptr(dimg),
             _s(dy))


def pack_first_dgrad_weights(w: torch.Tensor) -> torch.Tensor:
    """conv1_1 weights ``[64, 3, 3, 3]`` -> ``[9, 16, 64]`` for the tensor-core input gradient
    (flipped/transposed packing, 3 real rows padded to the smallest UMMA N)."""
    _wf, wd = pack_conv_weights(w)          # wd: [9, 3, 64], tf32-rounded
    out = torch.zeros(9, 16, w.shape[0], device=w.device, dtype=torch.float32)
    out[:, :3, :] = wd
    return out


def pack_first_dgrad_rows(w: torch.Tensor) -> torch.Tensor:
    """conv1_1 weights ``[64, 3, 3, 3]`` -> ``[3, 16, 64]`` for ``conv3x3_first_dgrad_rows``:
    ``out[t][kx * 3 + ci][co] = w[co][ci][2 - t][kx]`` (tf32-rounded), rows 9..15 zero."""
    _wf, wd = pack_conv_weights(w)          # wd[ky' * 3 + kx'][ci][co] = w[co][ci][2 - ky'][2 - kx']
    out = torch.zeros(3, 16, w.shape[0], device=w.device, dtype=torch.float32)
    for t in range(3):
        for kx in range(3):
            out[t, kx * 3:kx * 3 + 3, :] = wd[t * 3 + 2 - kx]
    return out


def conv3x3_first_dgrad_rows(dy: torch.Tensor, w_rows: torch.Tensor, dimg: torch.Tensor) -> None:
    """conv1_1 input gradient, NHWC ``dy`` -> NCHW ``dimg`` (x taps folded into the GEMM's N)."""
    _chk(dy, "dy")
    h, wd, cout = dy.shape
    nat.call("stv_conv3x3_first_dgrad_rows", nat.ptr(dy), nat.ptr(w_rows), h, wd, cout,
             nat.ptr(dimg), _s(dy))


def conv3x3_first_dgrad_tc(dy: torch.Tensor, w16: torch.Tensor, dimg: torch.Tensor) -> None:
    _chk(dy, "dy")
    h, wd, cout = dy.shape
    nat.call("stv_conv3x3_first_dgrad_tc", nat.ptr(dy), nat.ptr(w16), h, wd, cout, nat.ptr(dimg),
             _s(dy))


def maxpool2_fwd(x: torch.Tensor, y: torch.Tensor) -> None:
    _chk(x, "x")
    h, wd, c = x.shape
    nat.call("stv_maxpool2_fwd", nat.ptr(x), h, wd, c, nat.ptr(y), _s(x))


def maxpool2_bwd(dy: torch.Tensor, x: torch.Tensor, dx: torch.Tensor, *, relu_mask: bool) -> None:
    _chk(dy, "dy")
    h, wd, c = x.shape
    nat.call("stv_maxpool2_bwd", nat.ptr(dy), nat.ptr(x), h, wd, c, int(relu_mask), nat.ptr(dx),
             _s(x))


def relu_fwd(x: torch.Tensor, y: torch.Tensor) -> None:
    nat.call("stv_relu_fwd", nat.ptr(x), x.numel(), nat.ptr(y), _s(x))


def relu_fwd_bits(x: torch.Tensor, y: torch.Tensor, bits: torch.Tensor | None) -> None:
    """``y = tf32(relu(x))`` and its sign bits (``stv_relu_fwd_bits``)."""
    nat.call("stv_relu_fwd_bits", nat.ptr(x), x.numel(), nat.ptr(y), nat.ptr(bits), _s(x))


def relu_bwd(dy: torch.Tensor, x: torch.Tensor, dx: torch.Tensor, *, accumulate: bool) -> None:
    nat.call("stv_relu_bwd", nat.ptr(dy), nat.ptr(x), x.numel(), int(accumulate), nat.ptr(dx),
             _s(x))


def add_inplace(dst: torch.Tensor, src: torch.Tensor) -> None:
    nat.call("stv_add_inplace", nat.ptr(dst), nat.ptr(src), dst.numel(), _s(dst))


def gram_workspace(hw: int, channels: int, device: torch.device) -> torch.Tensor:
    nbytes = nat.gram_workspace_bytes(hw, channels)
    return torch.empty((nbytes + 3) // 4, device=device, dtype=torch.float32)


def gram_loss_fwd(x: torch.Tensor, workspace: torch.Tensor, *, target: torch.Tensor | None = None,
                  gram_out: torch.Tensor | None = None, s_out: torch.Tensor | None = None,
                  loss_out: torch.Tensor | None = None,
                  clamp_max: float = GRAM_MATRIX_CLAMP_MAX) -> None:
    """x: ``[H, W, C]`` (or ``[HW, C]``) features.  See ``stv_gram_loss_fwd``."""
    _chk(x, "x")
    c = x.shape[-1]
    hw = x.numel() // c
    nat.call("stv_gram_loss_fwd", nat.ptr(x), hw, c, nat.ptr(workspace), workspace.numel() * 4,
             nat.ptr(target), float(clamp_max), nat.ptr(gram_out), nat.ptr(s_out),
             nat.ptr(loss_out), _s(x))


def gram_partial_r(x: torch.Tensor, workspace: torch.Tensor, r_out: torch.Tensor) -> None:
    """Raw ``F_band F_band^T`` of this GPU's pixels (row-band sharding); see ``stv_gram_partial_r``."""
    _chk(x, "x")
    c = x.shape[-1]
    hw = x.numel() // c
    nat.call("stv_gram_partial_r", nat.ptr(x), hw, c, nat.ptr(workspace), workspace.numel() * 4,
             nat.ptr(r_out), _s(x))


def gram_from_r(r: torch.Tensor, n_total: float, scratch: torch.Tensor, *,
                target: torch.Tensor | None = None, gram_out: torch.Tensor | None = None,
                s_out: torch.Tensor | None = None, loss_out: torch.Tensor | None = None,
                clamp_max: float = GRAM_MATRIX_CLAMP_MAX) -> None:
    """Clamp / normalise / loss / backward seed from an all-reduced raw R."""
    c = r.shape[0]
    nat.call("stv_gram_from_r", nat.ptr(r), c, float(n_total), nat.ptr(target), float(clamp_max),
             nat.ptr(gram_out), nat.ptr(s_out), nat.ptr(loss_out), nat.ptr(scratch), _s(r))


def style_bwd(x: torch.Tensor, s: torch.Tensor, grad_w: torch.Tensor, dy: torch.Tensor, *,
              accumulate: bool) -> None:
    c = x.shape[-1]
    hw = x.numel() // c
    nat.call("stv_style_bwd", nat.ptr(x), nat.ptr(s), hw, c, nat.ptr(grad_w), int(accumulate),
             nat.ptr(dy), _s(x))


def content_loss_fwd(f: torch.Tensor, t: torch.Tensor, partials: torch.Tensor,
                     loss_out: torch.Tensor) -> None:
    nat.call("stv_content_loss_fwd", nat.ptr(f), nat.ptr(t), f.numel(), nat.ptr(partials),
             nat.ptr(loss_out), _s(f))


def content_loss_bwd(f: torch.Tensor, t: torch.Tensor, grad_w: torch.Tensor, df: torch.Tensor, *,
                     accumulate: bool) -> None:
    nat.call("stv_content_loss_bwd", nat.ptr(f), nat.ptr(t), f.numel(), nat.ptr(grad_w),
             int(accumulate), nat.ptr(df), _s(f))


def adam_step(x: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, beta1: float,
              beta2: float, eps: float, step_size: float, bias2_sqrt: float) -> None:
    nat.call("stv_adam_step", nat.ptr(x), nat.ptr(g), nat.ptr(m), nat.ptr(v), x.numel(), beta1,
             beta2, eps, step_size, bias2_sqrt, _s(x))


def adam_step_dev(x: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
                  state3: torch.Tensor, *, lr: float, beta1: float, beta2: float,
                  eps: float) -> None:
    nat.call("stv_adam_step_dev", nat.ptr(x), nat.ptr(g), nat.ptr(m), nat.ptr(v), x.numel(), lr,
             beta1, beta2, eps, nat.ptr(state3), _s(x))


def dot(a: torch.Tensor, b: torch.Tensor, partials: torch.Tensor, out: torch.Tensor) -> None:
    nat.call("stv_dot", nat.ptr(a), nat.ptr(b), a.numel(), nat.ptr(partials), nat.ptr(out), _s(a))


def absmax_sum(a: torch.Tensor, partials: torch.Tensor, out2: torch.Tensor) -> None:
    nat.call("stv_absmax_sum", nat.ptr(a), a.numel(), nat.ptr(partials), nat.ptr(out2), _s(a))


def axpy(alpha: float | torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> None:
    """y += alpha * x (alpha: host float or 0-dim/1-elem device tensor)."""
    if isinstance(alpha, torch.Tensor):
        nat.call("stv_axpy", nat.ptr(alpha), 0.0, nat.ptr(x), nat.ptr(y), x.numel(), _s(x))
    else:
        nat.call("stv_axpy", None, float(alpha), nat.ptr(x), nat.ptr(y), x.numel(), _s(x))


def scale(alpha: float | torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> None:
    """y = alpha * x."""
    if isinstance(alpha, torch.Tensor):
        nat.call("stv_scale", nat.ptr(alpha), 0.0, nat.ptr(x), nat.ptr(y), x.numel(), _s(x))
    else:
        nat.call("stv_scale", None, float(alpha), nat.ptr(x), nat.ptr(y), x.numel(), _s(x))


def lbfgs_workspace_floats(n: int, history: int) -> int:
    return int(nat.load().stv_lbfgs_workspace_floats(n, history))


def lbfgs_step(x: torch.Tensor, g: torch.Tensor, hist_s: torch.Tensor, hist_y: torch.Tensor,
               prev_g: torch.Tensor, d: torch.Tensor, workspace: torch.Tensor, *, history: int,
               lr: float, tolerance_grad: float, tolerance_change: float) -> None:
    """One device-resident L-BFGS step (see ``stv_lbfgs_step``)."""
    nat.call("stv_lbfgs_step", nat.ptr(x), nat.ptr(g), x.numel(), history, nat.ptr(hist_s),
             nat.ptr(hist_y), nat.ptr(prev_g), nat.ptr(d), nat.ptr(workspace), lr, tolerance_grad,
             tolerance_change, _s(x))


def frame_to_u8(img: torch.Tensor, out: torch.Tensor, *, denormalize: bool,
                rounding: bool = False) -> None:
    """img ``[1,3,H,W]`` fp32 -> out ``[H,W,3]`` uint8 (device)."""
    h, wd = img.shape[-2], img.shape[-1]
    nat.call("stv_frame_to_u8", nat.ptr(img), h, wd, int(denormalize), int(rounding),
             nat.ptr(out), _s(img))


def nchw_to_nhwc(src: torch.Tensor) -> torch.Tensor:
    """``[1,C,H,W]`` or ``[C,H,W]`` -> ``[H,W,C]`` (copy)."""
    c, h, wd = src.shape[-3], src.shape[-2], src.shape[-1]
    dst = torch.empty(h, wd, c, device=src.device, dtype=torch.float32)
    nat.call("stv_nchw_to_nhwc", nat.ptr(src), c, h, wd, nat.ptr(dst), _s(src))
    return dst


def nhwc_to_nchw(src: torch.Tensor) -> torch.Tensor:
    """``[H,W,C]`` -> ``[1,C,H,W]`` (copy)."""
    h, wd, c = src.shape
    dst = torch.empty(1, c, h, wd, device=src.device, dtype=torch.float32)
    nat.call("stv_nhwc_to_nchw", nat.ptr(src), c, h, wd, nat.ptr(dst), _s(src))
    return dst


def finite_flags(vals: torch.Tensor, flags: torch.Tensor) -> None:
    nat.call("stv_finite_flags", nat.ptr(vals), vals.numel(), nat.ptr(flags), _s(vals))


def halo_exchange(mine: torch.Tensor, *, up_ptr: int | None, down_ptr: int | None,  # noqa: PLR0913
                  rows: int, rows_up: int, rows_down: int, row_floats: int, planes: int,
                  flags_mine: torch.Tensor, flags_up_ptr: int | None, flags_down_ptr: int | None,
                  epoch: torch.Tensor, done: torch.Tensor, slot: int,
                  wait_ready: bool = True) -> None:
    """Push this rank's boundary rows into the neighbours' halo rows through peer-mapped memory and
    wait for theirs (``stv_halo_exchange``).  ``*_ptr`` are this process's mappings of the
    neighbours' buffers (None at the image boundary)."""
    nat.call("stv_halo_exchange", nat.ptr(mine), up_ptr, down_ptr, rows, rows_up, rows_down,
             row_floats, planes, nat.ptr(flags_mine), flags_up_ptr, flags_down_ptr, nat.ptr(epoch),
             nat.ptr(done), slot, int(wait_ready), _s(mine))


def step_scores(losses: torch.Tensor, n_style: int, n_content: int, style_w: float,  # noqa: PLR0913
                content_w: float, scores3: torch.Tensor, *, loss_ring: torch.Tensor | None = None,
                finite_ring: torch.Tensor | None = None,
                counter: torch.Tensor | None = None) -> None:
    """{style, content, total} of one step, appended to device rings at a device step counter
    (``stv_step_scores``)."""
    capacity = int(loss_ring.shape[0]) if loss_ring is not None else \
        (int(finite_ring.shape[0]) if finite_ring is not None else 0)
    nat.call("stv_step_scores", nat.ptr(losses), n_style, n_content, float(style_w),
             float(content_w), nat.ptr(scores3), nat.ptr(loss_ring), nat.ptr(finite_ring), capacity,
             nat.ptr(counter), _s(losses))


def conv_igemm2_ex(x: torch.Tensor, w_packed: torch.Tensor, *, taps: int,  # noqa: PLR0913
                   bias: torch.Tensor | None = None, alpha: torch.Tensor | None = None,
                   mask_src: torch.Tensor | None = None, add_src: torch.Tensor | None = None,
                   out_pre: torch.Tensor | None = None, out_post: torch.Tensor | None = None,
                   block_n: int = 0, m_halves: int = 0, tw: int = 0) -> None:
    """Test hook: the persistent tap-reusing conv with explicit tile selection."""
    h, wd, c = x.shape
    n = w_packed.shape[-2]
    nat.call("stv_conv_igemm2_ex", nat.ptr(x), nat.ptr(w_packed), h, wd, c, n, taps,
             nat.ptr(bias), nat.ptr(alpha), nat.ptr(mask_src), nat.ptr(add_src), nat.ptr(out_pre),
             nat.ptr(out_post), block_n, m_halves, tw, _s(x))


def conv_ref(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor | None, *, taps: int,
             relu: bool) -> torch.Tensor:
    """Test hook: naive CUDA-core conv with the same packed weights."""
    h, wd, c = x.shape
    n = w_packed.shape[-2]
    out = torch.empty(h, wd, n, device=x.device, dtype=torch.float32)
    nat.call("stv_conv_ref", nat.ptr(x), nat.ptr(w_packed), nat.ptr(bias), h, wd, c, n, taps,
             int(relu), nat.ptr(out), _s(x))
    return out


def conv_set_tuning(pair_mode: int = -1, a_stages: int = 0, b_stages: int = 0, tps: int = 0) -> None:
    """Process-wide conv tuning knobs (see ``stv_conv_set_tuning``); defaults restore the rule table."""
    nat.call("stv_conv_set_tuning", int(pair_mode), int(a_stages), int(b_stages), int(tps))


def conv_set_epilogue(staged_mode: int = -1) -> None:
    """-1: measured rule, 0: direct stores, 1: coalesced (staged) stores wherever the tile allows."""
    nat.call("stv_conv_set_epilogue", int(staged_mode))


def conv_set_split(mode: int = -1) -> None:
    """-1: built-in rule, 0: single issuer for one-half tiles, 1: split-K second issuer wherever legal."""
    nat.call("stv_conv_set_split", int(mode))


def conv_set_resident(mode: int = -1) -> None:
    """-1: built-in rule, 0: never keep the weights of 64 -> 64 layers resident (A/B runs)."""
    nat.call("stv_conv_set_resident", int(mode))


def conv_set_pool_smem(mode: int = -1) -> None:
    """-1: built-in rule, 0: fused pool by warp shuffles only (A/B runs)."""
    nat.call("stv_conv_set_pool_smem", int(mode))


def conv_plan_override(h: int = 0, w: int = 0, c: int = 0, n: int = 0, *, backward: bool = False,  # noqa: PLR0913
                       block_n: int = 0, m_halves: int = 0, pair: int = -1, depth: int = 0,
                       tps: int = 0) -> None:
    """Sweep hook: force the tile plan of one layer shape (``h <= 0`` clears all overrides)."""
    nat.call("stv_conv_plan_override", h, w, c, n, int(backward), block_n, m_halves, pair, depth, tps)


def conv_set_pair_mode(mode: int) -> None:
    """-1: built-in rule table, 0: single-CTA conv tiles only, 1: CTA pairs wherever legal."""
    conv_set_tuning(mode)


def image_from_u8(img_hwc: torch.Tensor, *, normalize: bool) -> torch.Tensor:
    """uint8 ``[H, W, 3]`` on the device -> fp32 ``[1, 3, H, W]`` (ToTensor + optional Normalize)."""
    if img_hwc.dtype != torch.uint8 or img_hwc.dim() != 3 or img_hwc.shape[2] != 3 \
            or not img_hwc.is_contiguous():
        msg = "image_from_u8 expects a contiguous uint8 [H, W, 3] tensor"
        raise ValueError(msg)
    h, w, _ = img_hwc.shape
    out = torch.empty(1, 3, h, w, device=img_hwc.device, dtype=torch.float32)
    nat.call("stv_image_from_u8", nat.ptr(img_hwc), h, w, int(normalize), nat.ptr(out), _s(img_hwc))
    return out
