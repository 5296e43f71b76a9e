"""Internal constants of the hot path (same values as the reference's constants.py /
config_defaults.py, which are part of its public behaviour)."""
from __future__ import annotations

from typing import Literal

# reference constants.py:11-15
IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]
GRAM_MATRIX_CLAMP_MAX = 5e5
DENORM_VIEW_SHAPE = (1, 3, 1, 1)
CSV_LOGGING_RECOMMENDED_STEPS = 2000
MIN_DIMENSION = 64
MAX_DIMENSION = 3000
VIDEO_QUALITY_MIN = 1
VIDEO_QUALITY_MAX = 10

# reference type_defs.py
InitMethod = Literal["content", "random", "white"]
VideoMode = Literal["realtime", "postprocess"]
LossHistory = dict[str, list[float]]

# reference config_defaults.py (user-facing defaults)
DEFAULT_STEPS = 1500
DEFAULT_STYLE_WEIGHT = 1e5
DEFAULT_CONTENT_WEIGHT = 1.0
DEFAULT_LEARNING_RATE = 1.0
DEFAULT_INIT_METHOD: InitMethod = "random"
DEFAULT_SEED = 0
DEFAULT_NORMALIZE = True
DEFAULT_LBFGS_MAX_ITER = 1
DEFAULT_LBFGS_MAX_EVAL = 1
DEFAULT_STYLE_LAYERS: tuple[int, ...] = (0, 5, 10, 19, 28)
DEFAULT_CONTENT_LAYERS: tuple[int, ...] = (21,)
DEFAULT_SAVE_EVERY = 20
DEFAULT_FPS = 10
DEFAULT_VIDEO_QUALITY = 10
DEFAULT_CREATE_VIDEO = True
DEFAULT_FINAL_ONLY = False
DEFAULT_VIDEO_INTRO_ENABLED = True
DEFAULT_VIDEO_INTRO_DURATION = 10.0
DEFAULT_VIDEO_OUTRO_DURATION = 10.0
DEFAULT_VIDEO_FINAL_FRAME_COMPARE = True
DEFAULT_VIDEO_MODE: VideoMode = "realtime"
DEFAULT_CREATE_GIF = False
DEFAULT_GIF_INCLUDE_INTRO = False
DEFAULT_GIF_INCLUDE_OUTRO = False
DEFAULT_DEVICE = "cuda"
DEFAULT_LOG_EVERY = 10
DEFAULT_OUTPUT_DIR = "out"
INTRO_MAX_CROSSFADE_FRAMES = 12  # reference video.py:74
