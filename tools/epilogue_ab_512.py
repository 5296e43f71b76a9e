"""Direct vs coalesced epilogue stores on the 512x512 layer shapes (single-wave launches: the
epilogue of a CTA's only tile is not hidden behind the next tile's MMAs).

    python tools/epilogue_ab_512.py
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from style_transfer_visualizer_b200 import ops  # noqa: E402
from tools.epilogue_ab import timed  # noqa: E402


def main() -> None:
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(512, 512, 64, 64), (256, 256, 64, 128), (256, 256, 128, 128), (128, 128, 128, 256),
              (128, 128, 256, 256), (64, 64, 256, 512), (64, 64, 512, 512), (32, 32, 512, 512)]
    for (h, w, cin, cout) in shapes:
        x = torch.randn(h, w, cin, device=dev, generator=g).relu()
        wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        wf, wd = ops.pack_conv_weights(wt)
        b = torch.zeros(cout, device=dev)
        pre = torch.empty(h, w, cout, device=dev)
        post = torch.empty(h, w, cout, device=dev)
        bits = ops.relu_bits_buffer(h, w, cout, dev)
        dy = torch.randn(h, w, cout, device=dev, generator=g)
        dx = torch.zeros(h, w, cin, device=dev)
        abits = ops.relu_bits_buffer(h, w, cin, dev)
        abits.random_(0, 2 ** 31 - 1)
        gf = 2.0 * 9 * cin * cout * h * w / 1e9
        rows = [
            ("fwd post+bits", lambda: ops.conv3x3_fwd(x, wf, b, None, post, out_bits=bits)),
            ("fwd pre+post+bits", lambda: ops.conv3x3_fwd(x, wf, b, pre, post, out_bits=bits)),
            ("fwd pre only", lambda: ops.conv3x3_fwd(x, wf, b, pre, None)),
            ("dgrad bits", lambda: ops.conv3x3_dgrad(dy, wd, dx, relu_bits=abits)),
            ("dgrad bits+acc", lambda: ops.conv3x3_dgrad(dy, wd, dx, relu_bits=abits, accumulate=True)),
        ]
        print(f"--- {h}x{w} {cin}->{cout}  ({gf:.1f} GFLOP)", flush=True)
        for name, fn in rows:
            out = []
            for mode in (0, 1):
                ops.conv_set_epilogue(mode)
                out.append(timed(fn, reps=20))
            ops.conv_set_epilogue(-1)
            rule = timed(fn, reps=20)
            print(f"   {name:<22s} direct {out[0]:7.1f} us  staged {out[1]:7.1f} us  rule {rule:7.1f} us "
                  f"({gf / rule * 1e3:6.1f} TF/s)", flush=True)
        # forced tile families for the wide layers: is a narrower, coalesced-store tile better than
        # the 256-wide direct-store one when every CTA has a single tile?
        if cout >= 256:
            for (bn, mh, pair) in [(256, 1, 1), (128, 2, 1), (128, 1, 0), (128, 1, 1)]:
                for mode in (0, 1):
                    ops.conv_set_tuning(pair)
                    ops.conv_set_epilogue(mode)
                    try:
                        us = timed(lambda: ops.conv_igemm2_ex(x, wf, taps=9, bias=b, out_pre=pre,  # noqa: B023
                                                              block_n=bn, m_halves=mh), reps=20)  # noqa: B023
                        print(f"   forced bn{bn} mh{mh} pair{pair} staged{mode} (out_pre only) "
                              f"{us:7.1f} us ({gf / us * 1e3:6.1f} TF/s)", flush=True)
                    except Exception as exc:  # noqa: BLE001
                        print(f"   forced bn{bn} mh{mh} pair{pair} staged{mode}: {exc}", flush=True)
            ops.conv_set_tuning()
            ops.conv_set_epilogue(-1)
    ops.conv_set_epilogue(-1)


if __name__ == "__main__":
    main()
