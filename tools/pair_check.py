"""CTA-pair (cta_group::2) conv vs the single-CTA kernel: exactness and speed per layer shape.

    python tools/pair_check.py [--big]     (one process; run under `timeout`)
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from style_transfer_visualizer_b200 import ops  # noqa: E402


def time_ms(fn, iters: int = 10) -> float:  # noqa: ANN001
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main() -> None:
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    big = "--big" in sys.argv
    small = [(40, 56, 64, 64, 9), (33, 47, 128, 256, 9), (24, 40, 256, 128, 9), (19, 21, 64, 128, 9),
             (16, 16, 512, 512, 9), (37, 29, 128, 128, 1), (8, 8, 64, 64, 9)]
    layers = [(1080, 1920, 64, 64, 9), (540, 960, 64, 128, 9), (540, 960, 128, 128, 9),
              (270, 480, 128, 256, 9), (270, 480, 256, 256, 9), (135, 240, 256, 512, 9),
              (135, 240, 512, 512, 9), (67, 120, 512, 512, 9),
              (512, 512, 64, 64, 9), (256, 256, 128, 128, 9), (128, 128, 256, 256, 9),
              (64, 64, 512, 512, 9), (32, 32, 512, 512, 9)]
    bad = 0
    for h, w, c, n, taps in (layers if big else small):
        x = torch.randn(h, w, c, device=dev, generator=g)
        wt = torch.randn(n, c, 3, 3, device=dev, generator=g) * 0.05
        wf, _ = ops.pack_conv_weights(wt)
        wp = wf if taps == 9 else wf[4:5].contiguous()
        bias = torch.randn(n, device=dev, generator=g)
        mask = torch.randn(h, w, n, device=dev, generator=g)
        add = torch.randn(h, w, n, device=dev, generator=g)
        flops = 2.0 * h * w * c * n * taps
        for bn in (256, 128, 64):
            if n % bn:
                continue
            for mh in (1, 2):
                res = {}
                for heavy in (False, True):
                    outs = []
                    for pair in (0, 1):
                        ops.conv_set_pair_mode(pair)
                        pre = torch.full((h, w, n), float("nan"), device=dev)
                        post = torch.full((h, w, n), float("nan"), device=dev)
                        kw = dict(taps=taps, bias=bias, out_pre=pre, out_post=post, block_n=bn,
                                  m_halves=mh)
                        if heavy:
                            kw.update(mask_src=mask, add_src=add)
                        ops.conv_igemm2_ex(x, wp, **kw)
                        torch.cuda.synchronize()
                        outs.append((pre, post))
                        if big and not heavy:
                            res[pair] = time_ms(lambda: ops.conv_igemm2_ex(x, wp, **kw))
                    same = torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
                    finite = bool(torch.isfinite(outs[1][0]).all())
                    if not (same and finite):
                        bad += 1
                        d = (outs[0][0] - outs[1][0]).abs().max().item()
                        print(f"MISMATCH {h}x{w} C{c} N{n} taps{taps} bn{bn} mh{mh} heavy={heavy} "
                              f"maxdiff={d} finite={finite}")
                if big:
                    print(f"{h}x{w} C{c}->N{n} bn{bn} mh{mh}: single {res[0]*1e3:7.1f} us "
                          f"({flops/res[0]/1e9:6.1f} TF/s)  pair {res[1]*1e3:7.1f} us "
                          f"({flops/res[1]/1e9:6.1f} TF/s)", flush=True)
                else:
                    print(f"ok {h}x{w} C{c} N{n} taps{taps} bn{bn} mh{mh}", flush=True)
    ops.conv_set_pair_mode(-1)
    print("PAIR CHECK", "FAILED" if bad else "PASSED", bad)


if __name__ == "__main__":
    main()
