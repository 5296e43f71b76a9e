"""Per-layer tile-plan sweep measured on the WHOLE graph-replayed step.

For every conv layer shape of the step (forward and input-gradient launch separately) each
candidate plan (N tile, pixel halves, CTA pair, ring depth) replaces the rule table's choice for
that one layer (``stv_conv_plan_override``); the step is re-captured and timed.  Unlike a per-launch
micro-benchmark this sees what the step sees: warm L2, programmatic dependent launch overlap, the
loss kernels on the side stream.

    python tools/plan_sweep.py --size 512 [--reps 200]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402


def build(h: int, w: int):  # noqa: ANN201
    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import synthetic
    from style_transfer_visualizer_b200.optim import FusedAdam

    dev = torch.device("cuda:0")
    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
    finally:
        cm.initialize_vgg = original
    content = synthetic.synthetic_image(1, h, w).to(dev)
    style = synthetic.synthetic_image(2, h, w).to(dev)
    model.set_targets(style, content)
    x = cm.initialize_input(content, "content")
    return model, x, FusedAdam([x], lr=0.01)


def time_step(model, x, opt, reps: int) -> float:  # noqa: ANN001
    """Median-of-3 microseconds per graph-replayed step with the current overrides."""
    from style_transfer_visualizer_b200.fused_step import FusedStep

    fs = FusedStep(model, x, opt, 1e5, 1.0)
    for _ in range(10):
        fs.step()
    torch.cuda.synchronize()
    out = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fs.step()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) * 1e3 / reps)
    return sorted(out)[1]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="512")
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--layers", default="", help="comma list of layer indices to sweep (default all)")
    ap.add_argument("--ab-split", action="store_true",
                    help="only compare the split-K second issuer: off / rule / everywhere legal")
    ap.add_argument("--ab-pool", action="store_true", help="only compare the two fused-pool epilogues")
    ap.add_argument("--ab-resident", action="store_true",
                    help="only compare weight-stationary 64->64 layers: off / rule")
    args = ap.parse_args()
    from style_transfer_visualizer_b200 import ops

    h, w = (1080, 1920) if args.size == "1080p" else (int(args.size), int(args.size))
    model, x, opt = build(h, w)
    if args.ab_resident:
        for rnd in range(2):
            for mode, name in ((0, "weights streamed"), (-1, "rule (64->64 weight-stationary)")):
                ops.conv_set_resident(mode)
                us = time_step(model, x, opt, args.reps)
                print(f"round {rnd} {name:32s}: {us:8.1f} us/step ({1e6 / us:6.1f} steps/s)", flush=True)
        ops.conv_set_resident()
        return
    if args.ab_pool:
        for rnd in range(2):
            for mode, name in ((0, "pool by shuffles"), (-1, "rule (pool through the staging tile)")):
                ops.conv_set_pool_smem(mode)
                us = time_step(model, x, opt, args.reps)
                print(f"round {rnd} {name:36s}: {us:8.1f} us/step ({1e6 / us:6.1f} steps/s)", flush=True)
        ops.conv_set_pool_smem()
        return
    if args.ab_split:
        for rnd in range(2):
            for mode, name in ((0, "single issuer"), (-1, "rule"), (1, "split wherever legal")):
                ops.conv_set_split(mode)
                us = time_step(model, x, opt, args.reps)
                print(f"round {rnd} {name:22s}: {us:8.1f} us/step ({1e6 / us:6.1f} steps/s)", flush=True)
        ops.conv_set_split()
        return
    # VGG19 conv layers up to conv5_1: (cin, cout) and pooling before
    chans = [(64, 64), "P", (64, 128), (128, 128), "P", (128, 256), (256, 256), (256, 256), (256, 256),
             "P", (256, 512), (512, 512), (512, 512), (512, 512), "P", (512, 512)]
    shapes = []
    hh, ww = h, w
    for c in chans:
        if c == "P":
            hh, ww = hh // 2, ww // 2
            continue
        shapes.append((hh, ww, c[0], c[1]))
    base = time_step(model, x, opt, args.reps)
    print(f"rule table: {base:8.1f} us/step  ({1e6 / base:6.1f} steps/s)", flush=True)
    base2 = time_step(model, x, opt, args.reps)
    print(f"rule table again: {base2:8.1f} us/step (noise {abs(base2 - base):.1f} us)", flush=True)
    base = min(base, base2)
    want = {int(t) for t in args.layers.split(",") if t} if args.layers else None
    best_plans = []
    for li, (lh, lw, cin, cout) in enumerate(shapes):
        if want is not None and li not in want:
            continue
        for backward in (False, True):
            # the launch's GEMM: forward cin -> cout; input gradient cout -> cin
            c, n = (cout, cin) if backward else (cin, cout)
            if backward and n < 64:
                continue
            cands = []
            for bn in (256, 128, 64):
                if n % bn:
                    continue
                for mh in (1, 2):
                    for pair in (0, 1):
                        depths = (0,) if bn == 256 else (0, 4)
                        for depth in depths:
                            cands.append((bn, mh, pair, depth))
            rows = []
            for (bn, mh, pair, depth) in cands:
                ops.conv_plan_override()  # clear
                ops.conv_plan_override(lh, lw, c, n, backward=backward, block_n=bn, m_halves=mh,
                                       pair=pair, depth=depth)
                try:
                    us = time_step(model, x, opt, args.reps)
                except Exception as e:  # noqa: BLE001
                    print(f"   L{li} {'bwd' if backward else 'fwd'} bn{bn} mh{mh} pair{pair} d{depth}: "
                          f"fail {str(e)[:80]}", flush=True)
                    torch.cuda.synchronize()
                    continue
                rows.append((us, bn, mh, pair, depth))
            ops.conv_plan_override()
            rows.sort()
            tag = f"L{li} {lh}x{lw} {'dgrad' if backward else 'fwd'} {c}->{n}"
            print(f"{tag}:", flush=True)
            for us, bn, mh, pair, depth in rows[:4]:
                print(f"      bn{bn} mh{mh} pair{pair} depth{depth}: {us:8.1f} us/step  "
                      f"({base - us:+6.1f} us vs rule)", flush=True)
            if rows and base - rows[0][0] > 3.0:
                best_plans.append((lh, lw, c, n, backward, *rows[0][1:], base - rows[0][0]))
    print("---- all winning overrides together", flush=True)
    ops.conv_plan_override()
    for (lh, lw, c, n, backward, bn, mh, pair, depth, gain) in best_plans:
        print(f"   {lh}x{lw} {c}->{n} {'dgrad' if backward else 'fwd'}: bn{bn} mh{mh} pair{pair} "
              f"depth{depth}  (-{gain:.1f} us)")
        ops.conv_plan_override(lh, lw, c, n, backward=backward, block_n=bn, m_halves=mh, pair=pair,
                               depth=depth)
    us = time_step(model, x, opt, args.reps)
    print(f"combined: {us:8.1f} us/step ({1e6 / us:6.1f} steps/s) vs rule {base:8.1f}", flush=True)
    ops.conv_plan_override()


if __name__ == "__main__":
    main()
