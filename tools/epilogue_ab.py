"""Hot per-layer timings of the conv epilogue variants at the 1080p shapes of the 64- and 128-wide
layers (run on the GPU box): which variant of each forward / dgrad launch is the fast one.

    python tools/epilogue_ab.py
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from style_transfer_visualizer_b200 import ops  # noqa: E402


def timed(fn, reps: int = 10) -> float:  # noqa: ANN001
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main() -> None:
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    for (h, w, cin, cout) in [(1080, 1920, 64, 64), (540, 960, 64, 128), (540, 960, 128, 128),
                              (270, 480, 128, 256)]:
        x = torch.randn(h, w, cin, device=dev, generator=g).relu()
        wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        wf, wd = ops.pack_conv_weights(wt)
        b = torch.zeros(cout, device=dev)
        pre = torch.empty(h, w, cout, device=dev)
        post = torch.empty(h, w, cout, device=dev)
        pool = torch.empty(h // 2, w // 2, cout, device=dev)
        bits = ops.relu_bits_buffer(h, w, cout, dev)
        code = ops.pool_code_buffer(h, w, cout, dev)
        gf = 2.0 * 9 * cin * cout * h * w / 1e9
        rows = [
            ("fwd post", lambda: ops.conv3x3_fwd(x, wf, b, None, post)),
            ("fwd post+bits", lambda: ops.conv3x3_fwd(x, wf, b, None, post, out_bits=bits)),
            ("fwd pre+post+bits", lambda: ops.conv3x3_fwd(x, wf, b, pre, post, out_bits=bits)),
            ("fwd post+pool", lambda: ops.conv3x3_fwd(x, wf, b, None, post, out_pool=pool)),
            ("fwd pool+route", lambda: ops.conv3x3_fwd(x, wf, b, None, None, out_pool=pool,
                                                       out_code=code)),
        ]
        # dgrad of this conv: dy [h, w, cout] -> dx [h, w, cin]
        dy = torch.randn(h, w, cout, device=dev, generator=g)
        dx = torch.zeros(h, w, cin, device=dev)
        act = x
        abits = ops.relu_bits_buffer(h, w, cin, dev)
        ops.conv3x3_fwd(torch.randn(h, w, 64, device=dev, generator=g),
                        ops.pack_conv_weights(torch.randn(cin, 64, 3, 3, device=dev, generator=g))[0],
                        None, None, torch.empty(h, w, cin, device=dev), out_bits=abits)
        rows += [
            ("dgrad plain", lambda: ops.conv3x3_dgrad(dy, wd, dx)),
            ("dgrad relu_src", lambda: ops.conv3x3_dgrad(dy, wd, dx, relu_src=act)),
            ("dgrad bits", lambda: ops.conv3x3_dgrad(dy, wd, dx, relu_bits=abits)),
            ("dgrad relu_src+acc", lambda: ops.conv3x3_dgrad(dy, wd, dx, relu_src=act,
                                                             accumulate=True)),
            ("dgrad bits+acc", lambda: ops.conv3x3_dgrad(dy, wd, dx, relu_bits=abits,
                                                         accumulate=True)),
        ]
        if cin in (64, 128):
            feat = torch.randn(h, w, cin, device=dev, generator=g)
            sm = torch.randn(cin, cin, device=dev, generator=g) * 0.01
            sm = (sm + sm.t()) * 0.5
            gw = torch.tensor([1.0], device=dev)
            rows += [
                ("style_bwd alone", lambda: ops.style_bwd(feat, sm, gw, dx, accumulate=False)),
                ("dgrad bits+style fused", lambda: ops.conv3x3_dgrad_style(
                    dy, wd, dx, relu_bits=abits, feat=feat, s_mat=sm, grad_w=gw)),
            ]
        # un-pooling dgrad: output at 2h x 2w
        big = torch.zeros(2 * h, 2 * w, cin, device=dev)
        bcode = ops.pool_code_buffer(2 * h, 2 * w, cin, dev)
        bcode.random_(0, 2 ** 31 - 1)
        dpool = torch.empty(h, w, cin, device=dev)
        bigact = torch.randn(2 * h, 2 * w, cin, device=dev, generator=g).relu() if h <= 540 else None
        rows += [("dgrad unpool (writes 2h x 2w)", lambda: ops.conv3x3_dgrad_unpool(dy, wd, bcode, big))]
        if bigact is not None:
            rows += [("dgrad plain + maxpool2_bwd", lambda: (
                ops.conv3x3_dgrad(dy, wd, dpool),
                ops.maxpool2_bwd(dpool, bigact, big, relu_mask=True)))]
        print(f"--- {h}x{w} {cin}->{cout}  ({gf:.1f} GFLOP)", flush=True)
        for name, fn in rows:
            us = timed(fn)
            print(f"   {name:<32s} {us:8.1f} us   {gf / us * 1e3:7.1f} TF/s", flush=True)
        del x, pre, post, pool, dy, dx, big, bcode, bigact
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
