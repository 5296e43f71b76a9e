"""Model + loss construction with the reference's call surface, executed by sm_100a kernels.

Drop-in for the reference's ``core_model`` module (core_model.py:29-350): ``gram_matrix``,
``initialize_input``, ``initialize_vgg``, ``create_feature_blocks``, ``StyleContentModel`` and
``prepare_model_and_input`` keep their names, arguments, return types and error behaviour.  The
numerics are a ``torch.library`` custom op (``stv_b200::vgg_losses``) with an autograd formula
whose forward and backward run the kernel program in ``engine.py``; there is no PyTorch / CPU
fallback -- a non-sm_100 device raises.
"""
from __future__ import annotations

import weakref
from pathlib import Path
from urllib.parse import urlparse

import torch
from torch import nn

from . import _native as nat
from . import ops
from .config import OptimizationConfig
from .constants import GRAM_MATRIX_CLAMP_MAX, InitMethod
from .engine import VggLossEngine
from .logging_utils import logger

TensorList = list[torch.Tensor]

# ----------------------------------------------------------------------------------------------
# torch.library custom ops.  Engines are looked up by integer handle (ops take tensors and ints).
# ----------------------------------------------------------------------------------------------
# The registry holds engines WEAKLY: an engine (packed weights, side stream, multi-GB activation
# workspaces, tensors pinned by captured graphs) lives exactly as long as the model that owns it,
# like the reference's nn.Module state.
_ENGINES: "weakref.WeakValueDictionary[int, VggLossEngine]" = weakref.WeakValueDictionary()
_NEXT_HANDLE = [1]


def _register_engine(engine: VggLossEngine) -> int:
    handle = _NEXT_HANDLE[0]
    _NEXT_HANDLE[0] += 1
    _ENGINES[handle] = engine
    return handle


def _engine(handle: int) -> VggLossEngine:
    engine = _ENGINES.get(handle)
    if engine is None:
        msg = ("the StyleContentModel that produced this graph has been released (or moved to "
               "another device); run a new forward pass")
        raise RuntimeError(msg)
    return engine


@torch.library.custom_op("stv_b200::vgg_losses", mutates_args=(), device_types="cuda")
def vgg_losses(x: torch.Tensor, handle: int) -> torch.Tensor:
    """All style losses (ascending layer index) followed by all content losses, shape ``[n]``."""
    engine = _engine(handle)
    losses, _generation = engine.forward_losses(x)
    return losses.clone()


@vgg_losses.register_fake
def _(x: torch.Tensor, handle: int) -> torch.Tensor:
    engine = _engine(handle)
    return x.new_empty(len(engine.style_idx) + len(engine.content_idx))


@torch.library.custom_op("stv_b200::vgg_losses_backward", mutates_args=(), device_types="cuda")
def vgg_losses_backward(grad_losses: torch.Tensor, x: torch.Tensor, handle: int,
                        generation: int) -> torch.Tensor:
    """d(sum_k grad_losses[k] * loss_k)/dx, NCHW like ``x``.  Weights arrive on the device: the
    runner's ``style_w`` / ``content_w`` reach the kernels without a host sync."""
    engine = _engine(handle)
    grad = engine.backward_losses(int(x.shape[2]), int(x.shape[3]), grad_losses, generation)
    return grad.clone()


@vgg_losses_backward.register_fake
def _(grad_losses: torch.Tensor, x: torch.Tensor, handle: int, generation: int) -> torch.Tensor:
    return torch.empty_like(x)


def _setup_context(ctx, inputs, output) -> None:  # noqa: ANN001
    x, handle = inputs
    ctx.handle = handle
    # pair this forward with its activations: a later forward at the same size invalidates them
    ctx.generation = _engine(handle).generation_of(int(x.shape[2]), int(x.shape[3]))
    ctx.save_for_backward(x)


def _backward(ctx, grad_out: torch.Tensor):  # noqa: ANN001, ANN202
    (x,) = ctx.saved_tensors
    return vgg_losses_backward(grad_out.contiguous(), x, ctx.handle, ctx.generation), None


vgg_losses.register_autograd(_backward, setup_context=_setup_context)


# ----------------------------------------------------------------------------------------------
# reference-compatible functions
# ----------------------------------------------------------------------------------------------
def gram_matrix(tensor: torch.Tensor, clamp_max: float = GRAM_MATRIX_CLAMP_MAX) -> torch.Tensor:
    """Gram matrix of ``[batch, channels, h, w]`` features -> ``[batch*channels]^2`` (reference
    core_model.py:29-63: batch folded into channels, clamp before the 1/N).  Runs the tcgen05
    Gram kernel; forward only (the differentiable path is ``StyleContentModel``)."""
    b, c, h, w = tensor.size()
    nat.require_device(tensor.device)
    ch = b * c
    # The kernel takes 64 or a multiple of 128 channels; other counts are zero-padded (zero
    # channels add zero rows/columns to F F^T, which are sliced off again).  The 1/N uses the
    # padded count inside the kernel, so rescale by padded/true.
    padded = 64 if ch <= 64 else -(-ch // 128) * 128
    feats = tensor.detach().to(torch.float32).reshape(ch, h * w).t()  # [hw, ch]
    if padded != ch:
        feats = torch.nn.functional.pad(feats, (0, padded - ch))
    feats = feats.contiguous()
    work = ops.gram_workspace(h * w, padded, tensor.device)
    out = torch.empty(padded, padded, device=tensor.device, dtype=torch.float32)
    ops.gram_loss_fwd(feats, work, gram_out=out, clamp_max=clamp_max)
    if padded != ch:
        out = out[:ch, :ch] * (padded / ch)
    return out


def initialize_input(content_img: torch.Tensor, method: InitMethod) -> torch.Tensor:
    """Initial optimisation variable (reference core_model.py:66-100)."""
    if not isinstance(content_img, torch.Tensor):
        msg = f"Expected content_img to be a Tensor, got {type(content_img)}"
        raise TypeError(msg)
    if method == "content":
        start = content_img.clone()
    elif method == "random":
        start = torch.randn_like(content_img)
    elif method == "white":
        start = torch.ones_like(content_img)
    else:
        msg = f"Unsupported initialization method: {method}"
        raise ValueError(msg)
    return start.requires_grad_(True)  # noqa: FBT003


def initialize_vgg() -> nn.Module:
    """Pretrained, frozen, eval-mode ``vgg19().features`` (reference core_model.py:103-117).

    Offline machines (no weight download possible) can set ``STV_RANDOM_VGG_SEED=<int>`` to get the
    seeded random-init network the benchmarks and tests use; a warning is logged."""
    import os

    from torchvision.models import VGG19_Weights, vgg19

    seed = os.environ.get("STV_RANDOM_VGG_SEED")
    if seed is not None:
        from .synthetic import random_vgg19_features

        logger.warning("STV_RANDOM_VGG_SEED=%s: using RANDOM-INIT VGG19 weights (not pretrained)",
                       seed)
        return random_vgg19_features(int(seed))

    weights = VGG19_Weights.IMAGENET1K_V1
    cache = Path(torch.hub.get_dir()) / "checkpoints" / Path(urlparse(weights.url).path).name
    if cache.exists():
        logger.info("Using cached VGG19 weights at %s", cache)
    else:
        logger.info("Downloading VGG19 weights to %s", cache)
    features = vgg19(weights=weights).features.eval()
    for param in features.parameters():
        param.requires_grad_(False)  # noqa: FBT003
    return features


def create_feature_blocks(
    vgg: nn.Module,
    style_layers: list[int],
    content_layers: list[int],
) -> tuple[nn.ModuleList, list[int], list[int]]:
    """Cut the layer sequence after every tapped index (reference core_model.py:120-146).
    Returns ``(blocks, content_ids, style_ids)``; ReLUs become out-of-place; layers after the
    last tap are dropped.  Module names inside each block keep the original layer index."""
    blocks = nn.ModuleList()
    content_ids: list[int] = []
    style_ids: list[int] = []
    current = nn.Sequential()
    for idx, layer in enumerate(vgg.children()):
        current.add_module(str(idx), nn.ReLU(inplace=False) if isinstance(layer, nn.ReLU)
                           else layer)
        is_style, is_content = idx in style_layers, idx in content_layers
        if is_style or is_content:
            blocks.append(current)
            current = nn.Sequential()
            if is_style:
                style_ids.append(len(blocks) - 1)
            if is_content:
                content_ids.append(len(blocks) - 1)
    return blocks, content_ids, style_ids


class StyleContentModel(nn.Module):
    """VGG19 feature extractor + style/content losses (reference core_model.py:149-328).

    Public attributes match the reference: ``vgg_blocks``, ``style_ids``, ``content_ids``,
    ``style_targets`` (list of ``[C, C]`` Grams), ``content_targets`` (list of ``[1, C, h, w]``
    feature tensors).  ``forward(x)`` returns ``(style_losses, content_losses)`` -- lists of 0-dim
    tensors attached to the autograd graph of ``x``.
    """

    def __init__(self, style_layers: list[int], content_layers: list[int]) -> None:
        super().__init__()
        vgg = initialize_vgg()
        self.vgg_blocks, self.content_ids, self.style_ids = create_feature_blocks(
            vgg, style_layers, content_layers)
        self.style_targets: list[torch.Tensor] | None = None
        self.content_targets: list[torch.Tensor] | None = None
        self._engine: VggLossEngine | None = None
        self._handle: int | None = None

    # -- engine plumbing -------------------------------------------------------------------
    def _flat_layers(self) -> tuple[list[nn.Module], list[int], list[int]]:
        layers: list[nn.Module] = []
        style_idx: list[int] = []
        content_idx: list[int] = []
        for j, block in enumerate(self.vgg_blocks):
            layers.extend(block.children())
            last = len(layers) - 1
            if j in self.style_ids:
                style_idx.append(last)
            if j in self.content_ids:
                content_idx.append(last)
        return layers, style_idx, content_idx

    def engine_for(self, device: torch.device) -> VggLossEngine:
        """Build (once per device) the kernel program from the current block weights."""
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self._engine is None or self._engine.device != device:
            self.release()
            layers, style_idx, content_idx = self._flat_layers()
            self._engine = VggLossEngine(layers, style_idx, content_idx, device)
            self._handle = _register_engine(self._engine)
        return self._engine

    def release(self) -> None:
        """Drop the kernel program of this model (packed weights, workspaces, side stream); it is
        rebuilt on the next use.  Called automatically when the model moves or is collected."""
        engine, self._engine = self._engine, None
        if engine is not None:
            engine.release_workspaces()
            if self._handle is not None:
                _ENGINES.pop(self._handle, None)
        self._handle = None

    def _apply(self, fn, *args, **kwargs):  # noqa: ANN001, ANN002, ANN003, ANN202
        # .to(device) / .cuda() move the nn.Module weights; the packed copies are rebuilt lazily.
        self.release()
        return super()._apply(fn, *args, **kwargs)

    # -- reference API ---------------------------------------------------------------------
    def set_targets(self, style_img: torch.Tensor, content_img: torch.Tensor) -> None:
        """Precompute style Grams and content features (reference core_model.py:218-232)."""
        engine = self.engine_for(style_img.device)
        engine.compute_targets(style_img.detach(), content_img.detach())
        self.style_targets = list(engine.style_targets)
        # zero-copy NCHW-shaped views of the NHWC storage the kernels read
        self.content_targets = [t.permute(2, 0, 1).unsqueeze(0)
                                for t in engine.content_targets_nhwc]

    def _loss_vector(self, x: torch.Tensor) -> torch.Tensor:
        engine = self.engine_for(x.device)
        if self.style_targets is not None and engine.style_targets is None:
            msg = "model was moved to another device after set_targets(); call set_targets again"
            raise RuntimeError(msg)
        assert self._handle is not None
        return vgg_losses(x, self._handle)

    def _compute_style_losses(self, features: torch.Tensor, block_idx: int) -> torch.Tensor | None:
        """Style loss of one block from explicit NCHW features (reference core_model.py:234-264).
        Kept for API compatibility; ``forward`` computes all losses in one fused program."""
        if self.style_targets is None:
            msg = "style_targets must be set before computing losses."
            raise RuntimeError(msg)
        if block_idx not in self.style_ids:
            return None
        target = self.style_targets[self.style_ids.index(block_idx)]
        return torch.mean((gram_matrix(features) - target) ** 2)

    def _compute_content_losses(self, features: torch.Tensor,
                                block_idx: int) -> torch.Tensor | None:
        """Content loss of one block from explicit features (reference core_model.py:266-295)."""
        if self.content_targets is None:
            msg = "content_targets must be set before computing losses."
            raise RuntimeError(msg)
        if block_idx not in self.content_ids:
            return None
        target = self.content_targets[self.content_ids.index(block_idx)]
        return torch.mean((features - target) ** 2)

    def forward(self, x: torch.Tensor) -> tuple[TensorList, TensorList]:
        """Style and content losses of image ``x`` ``[1, 3, H, W]`` (reference
        core_model.py:297-328)."""
        if self.style_targets is None:
            msg = "style_targets must be set before computing losses."
            raise RuntimeError(msg)
        if self.content_targets is None:
            msg = "content_targets must be set before computing losses."
            raise RuntimeError(msg)
        vec = self._loss_vector(x)
        n_style = len(self.style_ids)
        n_content = len(self.content_ids)
        style_losses = [vec[k] for k in range(n_style)]
        content_losses = [vec[n_style + k] for k in range(n_content)]
        return style_losses, content_losses


def prepare_model_and_input(
    content_img: torch.Tensor,
    style_img: torch.Tensor,
    device: torch.device,
    optimization: OptimizationConfig,
) -> tuple[nn.Module, torch.Tensor, torch.optim.Optimizer]:
    """Create model, initial image and optimiser (reference core_model.py:331-350).

    The optimiser honours the reference default (L-BFGS with ``lr`` / ``lbfgs_max_iter`` /
    ``lbfgs_max_eval``); ``optimization.optimizer == "adam"`` selects the fused Adam update."""
    from .optim import FusedAdam, FusedLBFGS

    model = StyleContentModel(
        style_layers=optimization.style_layers,
        content_layers=optimization.content_layers,
    ).to(device)
    model.set_targets(style_img, content_img)
    input_img = initialize_input(content_img, optimization.init_method)
    if getattr(optimization, "optimizer", "lbfgs") == "adam":
        optimizer: torch.optim.Optimizer = FusedAdam([input_img], lr=optimization.lr)
    else:
        optimizer = FusedLBFGS(
            [input_img],
            lr=optimization.lr,
            max_iter=optimization.lbfgs_max_iter,
            max_eval=optimization.lbfgs_max_eval,
        )
    return model, input_img, optimizer
