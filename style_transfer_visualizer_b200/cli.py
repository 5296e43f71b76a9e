"""Command line front-end with the reference's flag names (cli.py:26-244) for the options that
reach the optimisation path; flags are optional overrides on top of ``--config`` TOML
(``argparse.SUPPRESS`` defaults so only user-supplied flags override the file)."""
from __future__ import annotations

import argparse
import sys

from . import main as stv_main
from .config import ConfigLoader, build_config_from_cli
from .logging_utils import logger

S = argparse.SUPPRESS


def build_arg_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="style-visualizer-b200",
                                description="Neural style transfer on B200 (sm_100a) kernels")
    p.add_argument("--content", required=True, help="path to the content image")
    p.add_argument("--style", required=True, help="path to the style image")
    p.add_argument("--config", help="TOML configuration file")
    out = p.add_argument_group("output")
    out.add_argument("--output", default=S)
    out.add_argument("--log-loss", dest="log_loss", default=S)
    out.add_argument("--log-every", dest="log_every", type=int, default=S)
    out.add_argument("--no-plot", dest="no_plot", action="store_true", default=S)
    opt = p.add_argument_group("optimization")
    opt.add_argument("--steps", type=int, default=S)
    opt.add_argument("--style-w", dest="style_w", type=float, default=S)
    opt.add_argument("--content-w", dest="content_w", type=float, default=S)
    opt.add_argument("--lr", type=float, default=S)
    opt.add_argument("--style-layers", dest="style_layers", default=S)
    opt.add_argument("--content-layers", dest="content_layers", default=S)
    opt.add_argument("--init-method", dest="init_method", choices=["content", "random", "white"],
                     default=S)
    opt.add_argument("--seed", type=int, default=S)
    opt.add_argument("--no-normalize", dest="no_normalize", action="store_true", default=S)
    opt.add_argument("--optimizer", choices=["lbfgs", "adam"], default=S,
                     help="lbfgs = reference default; adam = fused Adam update")
    vid = p.add_argument_group("video")
    vid.add_argument("--save-every", dest="save_every", type=int, default=S)
    vid.add_argument("--fps", type=int, default=S)
    vid.add_argument("--quality", type=int, default=S)
    vid.add_argument("--no-video", dest="no_video", action="store_true", default=S)
    vid.add_argument("--final-only", dest="final_only", action="store_true", default=S)
    vid.add_argument("--no-intro", dest="no_intro", action="store_true", default=S)
    hw = p.add_argument_group("hardware")
    hw.add_argument("--device", default=S)
    return p


def run_from_args(args: argparse.Namespace):  # noqa: ANN201
    values = vars(args)
    base = ConfigLoader.load(values["config"]) if values.get("config") else None
    cfg = build_config_from_cli(values, base_config=base)
    logger.info("Starting style transfer: content=%s style=%s steps=%d", args.content, args.style,
                cfg.optimization.steps)
    paths = stv_main.InputPaths(content_path=args.content, style_path=args.style)
    return stv_main.style_transfer(paths, cfg)


def main(argv: list[str] | None = None) -> int:
    args = build_arg_parser().parse_args(argv)
    run_from_args(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
