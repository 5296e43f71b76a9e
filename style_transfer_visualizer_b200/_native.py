"""ctypes binding of ``libstv_b200.so`` (the C-ABI declared in ``include/stv_b200.h``).

The shared library is built in-tree by ``build_native.py`` (``nvcc`` for sm_100a).  There is no
fallback: importing works on any machine (so CPU-only tests can check symbol coverage), but every
compute call requires a B200-class device and raises ``RuntimeError`` otherwise.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

import torch

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "lib" / "libstv_b200.so"

_vp = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_long
_f = ctypes.c_float
_sz = ctypes.c_size_t

class ConvDesc(ctypes.Structure):
    """``stv_conv_desc`` of include/stv_b200.h (field order and types must match)."""

    _fields_ = [
        ("x", _vp), ("w_packed", _vp),
        ("H", _i), ("W", _i), ("C", _i), ("N", _i), ("taps", _i),
        ("x_rows", _i), ("x_row0", _i),
        ("bias", _vp), ("alpha", _vp), ("mask_src", _vp), ("add_src", _vp),
        ("out_pre", _vp), ("out_post", _vp), ("round_flags", _i),
        ("out_pool", _vp), ("out_bits", _vp), ("out_code", _vp),
        ("mask_bits", _vp), ("unpool_code", _vp), ("H2", _i), ("W2", _i),
        ("style_x", _vp), ("style_s", _vp), ("style_alpha", _vp),
    ]


# name -> argtypes; restype is int unless listed in _RESTYPES
_SIGNATURES: dict[str, list] = {
    "stv_abi_version": [],
    "stv_device_check": [],
    "stv_pack_conv_weights": [_vp, _vp, _vp, _i, _i, _vp],
    "stv_conv3x3_first_fwd": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp],
    "stv_conv3x3_first_fwd_bits": [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "stv_conv3x3_first_fwd_tc": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "stv_conv3x3_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "stv_conv3x3_fwd_bits": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "stv_conv3x3_fwd_pool_code": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp],
    "stv_conv3x3_dgrad_bits": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "stv_conv3x3_dgrad_bits_style": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "stv_conv3x3_dgrad_unpool": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp],
    "stv_conv3x3_fwd_pool": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "stv_conv3x3_dgrad": [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "stv_conv3x3_first_dgrad": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "stv_conv3x3_first_dgrad_tc": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "stv_conv3x3_first_dgrad_rows": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "stv_maxpool2_fwd": [_vp, _i, _i, _i, _vp, _vp],
    "stv_maxpool2_bwd": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "stv_relu_fwd": [_vp, _l, _vp, _vp],
    "stv_relu_fwd_bits": [_vp, _l, _vp, _vp, _vp],
    "stv_relu_bwd": [_vp, _vp, _l, _i, _vp, _vp],
    "stv_add_inplace": [_vp, _vp, _l, _vp],
    "stv_gram_workspace_bytes": [_l, _i],
    "stv_gram_loss_fwd": [_vp, _l, _i, _vp, _sz, _vp, _f, _vp, _vp, _vp, _vp],
    "stv_gram_partial_r": [_vp, _l, _i, _vp, _sz, _vp, _vp],
    "stv_gram_from_r": [_vp, _i, ctypes.c_double, _vp, _f, _vp, _vp, _vp, _vp, _vp],
    "stv_style_bwd": [_vp, _vp, _l, _i, _vp, _i, _vp, _vp],
    "stv_reduce_scratch_floats": [],
    "stv_content_loss_fwd": [_vp, _vp, _l, _vp, _vp, _vp],
    "stv_content_loss_bwd": [_vp, _vp, _l, _vp, _i, _vp, _vp],
    "stv_adam_step": [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _f, _vp],
    "stv_adam_step_dev": [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _vp, _vp],
    "stv_dot": [_vp, _vp, _l, _vp, _vp, _vp],
    "stv_absmax_sum": [_vp, _l, _vp, _vp, _vp],
    "stv_axpy": [_vp, _f, _vp, _vp, _l, _vp],
    "stv_scale": [_vp, _f, _vp, _vp, _l, _vp],
    "stv_lbfgs_workspace_floats": [_l, _i],
    "stv_lbfgs_step": [_vp, _vp, _l, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp],
    "stv_frame_to_u8": [_vp, _i, _i, _i, _i, _vp, _vp],
    "stv_image_from_u8": [_vp, _i, _i, _i, _vp, _vp],
    "stv_nchw_to_nhwc": [_vp, _i, _i, _i, _vp, _vp],
    "stv_nhwc_to_nchw": [_vp, _i, _i, _i, _vp, _vp],
    "stv_finite_flags": [_vp, _i, _vp, _vp],
    "stv_conv3x3_desc": [ctypes.POINTER(ConvDesc), _vp],
    "stv_conv3x3_first_fwd_band": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "stv_halo_exchange": [_vp, _vp, _vp, _i, _i, _i, _l, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "stv_step_scores": [_vp, _i, _i, _f, _f, _vp, _vp, _vp, _i, _vp, _vp],
    "stv_conv_igemm2_ex": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                           _vp],
    "stv_conv_set_tuning": [_i, _i, _i, _i],
    "stv_conv_set_epilogue": [_i],
    "stv_conv_plan_override": [_i] * 10,
    "stv_conv_set_split": [_i],
    "stv_conv_set_resident": [_i],
    "stv_conv_set_pool_smem": [_i],
    "stv_conv_ref": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
}

_RESTYPES = {"stv_gram_workspace_bytes": _sz, "stv_lbfgs_workspace_floats": _sz}
EXPORTED_SYMBOLS = ("stv_last_error", *_SIGNATURES)

_lib: ctypes.CDLL | None = None


class NativeLibraryError(RuntimeError):
    """Raised when libstv_b200.so is missing or a native call reports an error."""


def load() -> ctypes.CDLL:
    """Load the shared library once and type its entry points."""
    global _lib  # noqa: PLW0603
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("STV_B200_LIB", LIB_PATH))
    if not path.exists():
        msg = (
            f"{path} not found: build it with `python build_native.py` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path."
        )
        raise NativeLibraryError(msg)
    lib = ctypes.CDLL(str(path))
    lib.stv_last_error.restype = ctypes.c_char_p
    lib.stv_last_error.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a contiguous tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        msg = "native kernels need contiguous tensors"
        raise ValueError(msg)
    return t.data_ptr()


def stream_ptr(device: torch.device | None = None) -> int:
    """cudaStream_t of torch's current stream (so launches are graph-capturable)."""
    return torch.cuda.current_stream(device).cuda_stream


# kernels enqueued per entry point (everything else launches exactly one)
_KERNELS_PER_CALL = {"stv_gram_loss_fwd": 3, "stv_gram_partial_r": 2, "stv_gram_from_r": 2, "stv_content_loss_fwd": 2, "stv_dot": 2,
                     "stv_absmax_sum": 2, "stv_adam_step_dev": 2, "stv_halo_exchange": 1, "stv_lbfgs_step": 5, "stv_lbfgs_workspace_floats": 0, "stv_abi_version": 0, "stv_conv_set_epilogue": 0, "stv_conv_plan_override": 0, "stv_conv_set_split": 0, "stv_conv_set_resident": 0, "stv_conv_set_pool_smem": 0, "stv_conv_set_tuning": 0,
                     "stv_device_check": 0}
_launches = [0]


def launch_count() -> int:
    """Kernels of this library enqueued so far by this process (graph replays not included)."""
    return _launches[0]


_call_device: list[torch.device | None] = [None]


def stream_for_call(device: torch.device) -> int:
    """``stream_ptr`` for the launch being assembled; also remembers ``device`` so that ``call``
    runs the launch with that GPU current.  The library uses the CURRENT device's context (kernel
    attributes, SM count) and the stream must belong to it: a tensor on ``cuda:1`` must not be
    launched on while ``cuda:0`` is current (the reference accepts ``--device cuda:1``)."""
    _call_device[0] = device
    return torch.cuda.current_stream(device).cuda_stream


def call_status(name: str, *args) -> int:  # noqa: ANN002
    """Like ``call`` but returns the status code 3 ("not fusable") instead of raising on it."""
    lib = load()
    _launches[0] += _KERNELS_PER_CALL.get(name, 1)
    dev, _call_device[0] = _call_device[0], None
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc not in (0, 3):
        detail = lib.stv_last_error().decode("utf-8", "replace")
        msg = f"{name} failed (status {rc}): {detail}"
        raise NativeLibraryError(msg)
    if rc == 3:
        _launches[0] -= _KERNELS_PER_CALL.get(name, 1)
    return rc


def call(name: str, *args) -> None:  # noqa: ANN002
    """Invoke an int-status entry point; raise with the library's message on failure."""
    lib = load()
    _launches[0] += _KERNELS_PER_CALL.get(name, 1)
    dev, _call_device[0] = _call_device[0], None
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        detail = lib.stv_last_error().decode("utf-8", "replace")
        msg = f"{name} failed (status {rc}): {detail}"
        raise NativeLibraryError(msg)


_device_ok: set[int] = set()


def require_device(device: torch.device) -> None:
    """Fail loudly unless ``device`` is an sm_100-class GPU with the library present."""
    if device.type != "cuda":
        msg = (
            "style_transfer_visualizer_b200 runs only on a CUDA sm_100a device "
            f"(got device '{device}'); there is no CPU fallback."
        )
        raise NativeLibraryError(msg)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _device_ok:
        return
    with torch.cuda.device(idx):
        call("stv_device_check")
    _device_ok.add(idx)


def gram_workspace_bytes(hw: int, channels: int) -> int:
    return int(load().stv_gram_workspace_bytes(hw, channels))


def reduce_scratch_floats() -> int:
    return int(load().stv_reduce_scratch_floats())
