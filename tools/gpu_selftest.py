"""Kernel-level self-test on a real B200: every C-ABI kernel against plain PyTorch fp32.

Each case runs in its own subprocess (own CUDA context, own timeout) so a trapping kernel cannot
take the remaining cases down.  Usage (on the GPU box):
    python tools/gpu_selftest.py [--cases a,b,...] [--out gpurun_out/selftest.log]
"""
from __future__ import annotations

import argparse
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = ["elementwise", "conv", "conv2", "conv_variants", "conv_compact", "halo", "gram", "perf"]


def rel_l2(a, b):  # noqa: ANN001, ANN201
    import torch

    a = a.double().flatten()
    b = b.double().flatten()
    return float(torch.linalg.vector_norm(a - b) / (torch.linalg.vector_norm(b) + 1e-300))


def report(name: str, err: float, tol: float, extra: str = "") -> bool:
    ok = err <= tol and err == err  # noqa: PLR0124
    print(f"{'PASS' if ok else 'FAIL'}  {name:<58s} err={err:.3e} tol={tol:.1e} {extra}", flush=True)
    return ok


def nhwc(t):  # noqa: ANN001, ANN201
    """[1,C,H,W] -> contiguous [H,W,C]."""
    return t[0].permute(1, 2, 0).contiguous()


def case_elementwise() -> bool:
    import torch
    import torch.nn.functional as F  # noqa: N812

    from style_transfer_visualizer_b200 import _native as nat
    from style_transfer_visualizer_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    nat.require_device(dev)
    ok = True
    g = torch.Generator(device="cuda").manual_seed(0)

    # relu + sign bits of a stored pre-activation (split ReLU of wide tapped layers)
    for n_pix, c in [(37 * 53, 256), (64 * 64, 512), (9, 32)]:
        xin = torch.randn(n_pix, c, device=dev, generator=g)
        y = torch.full((n_pix, c), float("nan"), device=dev)
        bits = torch.full((n_pix, c // 32), -1, device=dev, dtype=torch.int32)
        ops.relu_fwd_bits(xin, y, bits)
        want = xin.relu()
        sh = torch.arange(32, device=dev, dtype=torch.int32)
        unpacked = ((bits.unsqueeze(-1) >> sh) & 1).reshape(n_pix, c).bool()
        ok &= report(f"relu_fwd_bits values {n_pix}x{c}", rel_l2(y, want), 3e-4)
        ok &= report(f"relu_fwd_bits low bits zero {n_pix}x{c}",
                     float((y.view(torch.int32) & 0x1FFF).abs().max()), 0)
        ok &= report(f"relu_fwd_bits sign bits {n_pix}x{c}",
                     0.0 if torch.equal(unpacked, y > 0) else 1.0, 0.0)

    # pack weights
    w = torch.randn(128, 64, 3, 3, device=dev, generator=g)
    wf, wd = ops.pack_conv_weights(w)
    ok &= report("pack fwd (tf32-rounded)", rel_l2(wf, w.permute(2, 3, 0, 1).reshape(9, 128, 64)), 3e-4)
    ok &= report("pack fwd low bits zero", float((wf.view(torch.int32) & 0x1FFF).abs().max()), 0)
    ok &= report("pack dgrad (tf32-rounded)", rel_l2(wd, w.flip(2, 3).permute(2, 3, 1, 0).reshape(9, 64, 128)), 3e-4)

    for (h, wdt) in [(64, 64), (33, 47), (135, 250)]:
        # first conv fwd
        img = torch.randn(1, 3, h, wdt, device=dev, generator=g)
        w1 = torch.randn(64, 3, 3, 3, device=dev, generator=g) * 0.2
        b1 = torch.randn(64, device=dev, generator=g)
        pre = torch.empty(h, wdt, 64, device=dev)
        post = torch.empty_like(pre)
        ops.conv3x3_first_fwd(img, w1, b1, pre, post)
        ref = F.conv2d(img, w1, b1, padding=1)
        ok &= report(f"conv_first_fwd pre {h}x{wdt}", rel_l2(pre, nhwc(ref)), 2e-6)
        ok &= report(f"conv_first_fwd post {h}x{wdt} (tf32-rounded)", rel_l2(post, nhwc(ref.relu())), 3e-4)
        # first conv dgrad
        dy = torch.randn(1, 64, h, wdt, device=dev, generator=g)
        dimg = torch.empty(1, 3, h, wdt, device=dev)
        ops.conv3x3_first_dgrad(nhwc(dy), w1, dimg)
        refd = torch.nn.grad.conv2d_input(img.shape, w1, dy, padding=1)
        ok &= report(f"conv_first_dgrad {h}x{wdt}", rel_l2(dimg, refd), 2e-6)
        # pool fwd / bwd
        x = torch.randn(1, 64, h, wdt, device=dev, generator=g).relu().requires_grad_(True)
        y = F.max_pool2d(x, 2, 2)
        gy = torch.randn_like(y)
        y.backward(gy)
        yo = torch.empty(h // 2, wdt // 2, 64, device=dev)
        ops.maxpool2_fwd(nhwc(x.detach()), yo)
        ok &= report(f"maxpool fwd {h}x{wdt}", rel_l2(yo, nhwc(y.detach())), 0)
        dxo = torch.full((h, wdt, 64), 7.0, device=dev)
        ops.maxpool2_bwd(nhwc(gy), nhwc(x.detach()), dxo, relu_mask=False)
        ok &= report(f"maxpool bwd {h}x{wdt}", rel_l2(dxo, nhwc(x.grad)), 0)
        # fused relu gate: x is post-ReLU -> same as autograd through relu(max_pool)
        xr = torch.randn(1, 64, h, wdt, device=dev, generator=g).requires_grad_(True)
        yr = F.max_pool2d(xr.relu(), 2, 2)
        yr.backward(gy)
        ops.maxpool2_bwd(nhwc(gy), nhwc(xr.detach().relu()), dxo, relu_mask=True)
        ok &= report(f"maxpool+relu bwd {h}x{wdt}", rel_l2(dxo, nhwc(xr.grad)), 0)

    # relu
    x = torch.randn(4096 * 4, device=dev, generator=g)
    y = torch.empty_like(x)
    ops.relu_fwd(x, y)
    ok &= report("relu fwd", rel_l2(y, x.relu()), 0)
    dy = torch.randn_like(x)
    dx = torch.ones_like(x)
    ops.relu_bwd(dy, x, dx, accumulate=True)
    ok &= report("relu bwd acc (tf32-rounded)", rel_l2(dx, 1 + dy * (x > 0)), 3e-4)

    # content loss
    n = 512 * 33 * 47
    f = torch.randn(n, device=dev, generator=g)
    t = torch.randn(n, device=dev, generator=g)
    part = torch.empty(nat.reduce_scratch_floats() * 2, device=dev)
    loss = torch.zeros(1, device=dev)
    ops.content_loss_fwd(f, t, part, loss)
    ok &= report("content fwd", abs(float(loss) - float(F.mse_loss(f.double(), t.double())))
                 / float(F.mse_loss(f.double(), t.double())), 1e-6)
    gw = torch.tensor([3.0], device=dev)
    df = torch.empty_like(f)
    ops.content_loss_bwd(f, t, gw, df, accumulate=False)
    ok &= report("content bwd (tf32-rounded)", rel_l2(df, 3.0 * 2 * (f - t) / n), 3e-4)

    # adam vs torch.optim.Adam
    p = torch.randn(100003, device=dev, generator=g)
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=0.01)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    p_dev = p.clone()
    m2 = torch.zeros_like(p)
    v2 = torch.zeros_like(p)
    state = torch.zeros(3, device=dev)
    for step in range(1, 6):
        grad = torch.randn(p.shape, device=dev, generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        bc1 = 1 - 0.9 ** step
        bc2 = 1 - 0.999 ** step
        ops.adam_step(p, grad, m, v, beta1=0.9, beta2=0.999, eps=1e-8, step_size=0.01 / bc1,
                      bias2_sqrt=bc2 ** 0.5)
        ops.adam_step_dev(p_dev, grad, m2, v2, state, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8)
    ok &= report("adam 5 steps", rel_l2(p, p_ref.detach()), 1e-6)
    ok &= report("adam_dev 5 steps", rel_l2(p_dev, p_ref.detach()), 1e-6)

    # dot / absmax / axpy / scale
    a = torch.randn(1000003, device=dev, generator=g)
    b = torch.randn(1000003, device=dev, generator=g)
    out = torch.zeros(2, device=dev)
    ops.dot(a, b, part, out[:1])
    ok &= report("dot", abs(float(out[0]) - float(a.double() @ b.double())) / float(a.norm() * b.norm()), 1e-6)
    ops.absmax_sum(a, part, out)
    ok &= report("absmax", abs(float(out[0]) - float(a.abs().max())), 0)
    ok &= report("abssum", abs(float(out[1]) - float(a.double().abs().sum())) / float(a.double().abs().sum()), 1e-6)
    yv = b.clone()
    ops.axpy(0.5, a, yv)
    ok &= report("axpy host", rel_l2(yv, b + 0.5 * a), 1e-7)
    yv = b.clone()
    ops.axpy(torch.tensor([-2.0], device=dev), a, yv)
    ok &= report("axpy dev", rel_l2(yv, b - 2.0 * a), 1e-7)
    ops.scale(3.0, a, yv)
    ok &= report("scale", rel_l2(yv, 3.0 * a), 0)

    # frame conversion (bit exact vs the reference's torch+numpy arithmetic)
    for (h, wdt) in [(64, 64), (33, 47)]:
        img = torch.randn(1, 3, h, wdt, device=dev, generator=g) * 2
        img[0, 0, 0, 0] = float("nan")
        img[0, 1, 0, 1] = float("inf")
        img[0, 2, 0, 2] = float("-inf")
        for denorm in (True, False):
            mean = torch.tensor([0.485, 0.456, 0.406], device=dev).view(1, 3, 1, 1)
            std = torch.tensor([0.229, 0.224, 0.225], device=dev).view(1, 3, 1, 1)
            t_ = img * std + mean if denorm else img
            t_ = torch.nan_to_num(t_, nan=0.0, posinf=1.0, neginf=0.0).clamp(0, 1)
            ref_np = (t_.squeeze(0).permute(1, 2, 0).cpu().numpy() * 255).astype("uint8")
            out8 = torch.empty(h, wdt, 3, device=dev, dtype=torch.uint8)
            ops.frame_to_u8(img, out8, denormalize=denorm)
            diff = int((out8.cpu().numpy().astype(int) - ref_np.astype(int)).__abs__().max())
            ok &= report(f"frame_to_u8 {h}x{wdt} denorm={denorm}", float(diff), 0)

    # layout
    t4 = torch.randn(1, 64, 33, 47, device=dev, generator=g)
    ok &= report("nchw_to_nhwc", rel_l2(ops.nchw_to_nhwc(t4), nhwc(t4)), 0)
    ok &= report("nhwc_to_nchw", rel_l2(ops.nhwc_to_nchw(nhwc(t4)), t4), 0)
    vals = torch.tensor([1.0, float("nan"), float("inf"), 0.0], device=dev)
    flags = torch.zeros(4, device=dev, dtype=torch.int32)
    ops.finite_flags(vals, flags)
    ok &= report("finite_flags", float((flags.cpu() != torch.tensor([0, 1, 1, 0])).sum()), 0)
    torch.cuda.synchronize()
    return ok


def _conv_inputs(h, w, cin, cout, g, dev):  # noqa: ANN001, ANN202, PLR0913
    import torch

    x = torch.randn(1, cin, h, w, device=dev, generator=g)
    wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, device=dev, generator=g) * 0.1
    return x, wt, b


def case_conv() -> bool:
    import torch
    import torch.nn.functional as F  # noqa: N812

    from style_transfer_visualizer_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    ok = True
    # sanity of the on-device naive conv first (fp32 exact-ish)
    x, wt, b = _conv_inputs(20, 24, 64, 64, g, dev)
    wf, wd = ops.pack_conv_weights(wt)
    wt_rounded = wf.reshape(3, 3, 64, 64).permute(2, 3, 0, 1).contiguous()  # packing rounds to tf32
    ref = nhwc(F.conv2d(x, wt_rounded, b, padding=1))
    ok &= report("conv_ref vs torch", rel_l2(ops.conv_ref(nhwc(x), wf, b, taps=9, relu=False), ref), 5e-6)

    configs = [
        # h, w, cin, cout, block_n, th, tw
        (16, 16, 64, 64, 64, 8, 16),
        (16, 16, 32, 64, 64, 8, 16),
        (32, 32, 64, 128, 128, 8, 16),
        (32, 32, 128, 256, 256, 8, 16),
        (24, 40, 64, 64, 64, 4, 32),
        (24, 40, 64, 64, 64, 16, 8),
        (8, 128, 64, 64, 64, 1, 128),
        (10, 70, 64, 64, 64, 2, 64),
        (33, 47, 128, 128, 128, 8, 16),   # ragged edges
        (67, 120, 512, 512, 256, 16, 8),  # conv5_1 @1080p
        (67, 120, 512, 512, 128, 0, 0),   # auto patch
        (135, 240, 256, 512, 0, 0, 0),    # all auto
        (64, 64, 512, 512, 64, 0, 0),
    ]
    for (h, w, cin, cout, bn, th, tw) in configs:
        x, wt, b = _conv_inputs(h, w, cin, cout, g, dev)
        wf, wd = ops.pack_conv_weights(wt)
        ref = nhwc(F.conv2d(x, wt, b, padding=1))
        pre = torch.full((h, w, cout), float("nan"), device=dev)
        post = torch.full((h, w, cout), float("nan"), device=dev)
        ops.conv_igemm2_ex(nhwc(x), wf, taps=9, bias=b, out_pre=pre, out_post=post, block_n=bn,
                           tw=tw if tw in (8, 16, 32) else 0)
        torch.cuda.synchronize()
        tag = f"{h}x{w} {cin}->{cout} bn={bn} patch={th}x{tw}"
        ok &= report(f"igemm fwd pre  {tag}", rel_l2(pre, ref), 2e-3)
        ok &= report(f"igemm fwd post {tag}", rel_l2(post, ref.relu()), 2e-3)
        # dgrad through the same kernel with the flipped/transposed packing
        dy = torch.randn(1, cout, h, w, device=dev, generator=g)
        refd = nhwc(torch.nn.grad.conv2d_input(x.shape, wt, dy, padding=1))
        dx = torch.full((h, w, cin), float("nan"), device=dev)
        if cin % 64 == 0:
            ops.conv3x3_dgrad(nhwc(dy), wd, dx)
            torch.cuda.synchronize()
            ok &= report(f"igemm dgrad    {tag}", rel_l2(dx, refd), 2e-3)
    return ok


def case_conv2() -> bool:
    """Persistent tap-reusing kernel (conv_igemm2): forced tile shapes, auto selection, 1x1,
    epilogue options, and the N=16 first-layer input gradient."""
    import torch
    import torch.nn.functional as F  # noqa: N812

    from style_transfer_visualizer_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    ok = True
    configs = [
        # h, w, cin, cout, block_n, m_halves, tw
        (16, 16, 64, 64, 64, 1, 8),
        (16, 16, 64, 64, 64, 1, 16),
        (16, 16, 64, 64, 64, 1, 32),
        (32, 16, 64, 64, 64, 2, 8),
        (32, 32, 64, 128, 128, 2, 16),
        (40, 72, 128, 256, 256, 2, 8),
        (40, 72, 128, 256, 256, 1, 16),
        (33, 47, 128, 128, 128, 2, 8),     # ragged edges
        (33, 47, 64, 64, 64, 2, 32),
        (67, 120, 512, 512, 256, 2, 8),
        (67, 120, 512, 512, 0, 0, 0),      # auto
        (135, 240, 256, 512, 0, 0, 0),
        (64, 64, 512, 512, 0, 0, 0),
        (270, 480, 64, 64, 0, 0, 0),       # many tiles per CTA (persistent loop, acc double buffer)
        (200, 300, 128, 128, 0, 0, 0),
    ]
    for (h, w, cin, cout, bn, mh, tw) in configs:
        x, wt, b = _conv_inputs(h, w, cin, cout, g, dev)
        wf, wd = ops.pack_conv_weights(wt)
        ref = nhwc(F.conv2d(x, wt, b, padding=1))
        pre = torch.full((h, w, cout), float("nan"), device=dev)
        post = torch.full((h, w, cout), float("nan"), device=dev)
        ops.conv_igemm2_ex(nhwc(x), wf, taps=9, bias=b, out_pre=pre, out_post=post, block_n=bn,
                           m_halves=mh, tw=tw)
        torch.cuda.synchronize()
        tag = f"{h}x{w} {cin}->{cout} bn={bn} mh={mh} tw={tw}"
        ok &= report(f"igemm2 fwd pre  {tag}", rel_l2(pre, ref), 2e-3)
        ok &= report(f"igemm2 fwd post {tag}", rel_l2(post, ref.relu()), 2e-3)
    # hot-path entry points (auto tiles): fwd, dgrad with gate + accumulate, style bwd, first dgrad
    h, w, cin, cout = 37, 53, 128, 64
    x, wt, b = _conv_inputs(h, w, cin, cout, g, dev)
    wf, wd = ops.pack_conv_weights(wt)
    dy = torch.randn(1, cout, h, w, device=dev, generator=g)
    act = torch.randn(h, w, cin, device=dev, generator=g)
    prev = torch.randn(h, w, cin, device=dev, generator=g)
    refd = nhwc(torch.nn.grad.conv2d_input(x.shape, wt, dy, padding=1))
    dx = prev.clone()
    ops.conv3x3_dgrad(nhwc(dy), wd, dx, relu_src=act, accumulate=True)
    ok &= report("igemm2 dgrad + relu gate + accumulate", rel_l2(dx, refd * (act > 0) + prev), 2e-3)
    for (hh, ww) in [(64, 64), (33, 47), (135, 250), (8, 8), (5, 14), (9, 15), (17, 29), (272, 3840)]:
        w1 = torch.randn(64, 3, 3, 3, device=dev, generator=g) * 0.2
        dy1 = torch.randn(1, 64, hh, ww, device=dev, generator=g)
        ref1 = torch.nn.grad.conv2d_input((1, 3, hh, ww), w1, dy1, padding=1)
        dimg = torch.full((1, 3, hh, ww), float("nan"), device=dev)
        ops.conv3x3_first_dgrad_tc(nhwc(dy1), ops.pack_first_dgrad_weights(w1), dimg)
        ok &= report(f"first dgrad (tensor core, N=16) {hh}x{ww}", rel_l2(dimg, ref1), 2e-3)
        # x taps folded into N (the product path): same operands, same tolerance; every pixel written
        dimg2 = torch.full((1, 3, hh, ww), float("nan"), device=dev)
        ops.conv3x3_first_dgrad_rows(nhwc(dy1), ops.pack_first_dgrad_rows(w1), dimg2)
        ok &= report(f"first dgrad (x taps in N) {hh}x{ww}", rel_l2(dimg2, ref1), 2e-3)
        ok &= report(f"first dgrad (x taps in N) vs N=16 kernel {hh}x{ww}", rel_l2(dimg2, dimg), 2e-6)
    for c, hw in ((64, 33 * 47), (128, 64 * 64), (512, 67 * 120), (256, 1024 * 8 + 8)):
        xf = torch.randn(hw, c, device=dev, generator=g)
        sm = torch.randn(c, c, device=dev, generator=g)
        sm = (sm + sm.t()) * 0.5
        gw = torch.tensor([2.0], device=dev)
        out = torch.randn(hw, c, device=dev, generator=g)
        base = out.clone()
        ops.style_bwd(xf, sm, gw, out, accumulate=True)
        ok &= report(f"igemm2 style_bwd C={c} hw={hw}", rel_l2(out, base + 2.0 * (xf.double() @ sm.double()).float()), 2e-3)
    torch.cuda.synchronize()
    return ok


def case_conv_variants() -> bool:
    import torch
    import torch.nn.functional as F  # noqa: N812

    from style_transfer_visualizer_b200 import ops

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(2)
    ok = True
    h, w, cin, cout = 37, 53, 128, 64
    x, wt, b = _conv_inputs(h, w, cin, cout, g, dev)
    wf, wd = ops.pack_conv_weights(wt)
    dy = torch.randn(1, cout, h, w, device=dev, generator=g)
    act = torch.randn(h, w, cin, device=dev, generator=g)  # stand-in for the saved activation
    prev = torch.randn(h, w, cin, device=dev, generator=g)
    refd = nhwc(torch.nn.grad.conv2d_input(x.shape, wt, dy, padding=1))
    dx = prev.clone()
    ops.conv3x3_dgrad(nhwc(dy), wd, dx, relu_src=act, accumulate=True)
    ok &= report("dgrad + relu gate + accumulate", rel_l2(dx, refd * (act > 0) + prev), 2e-3)
    dx2 = torch.empty_like(prev)
    ops.conv3x3_dgrad(nhwc(dy), wd, dx2, relu_src=act)
    ok &= report("dgrad + relu gate", rel_l2(dx2, refd * (act > 0)), 2e-3)

    # 1x1 with alpha (style backward): dy = gw * X @ S
    for c in (64, 128, 256, 512):
        hw = 33 * 47
        xf = torch.randn(hw, c, device=dev, generator=g)
        s = torch.randn(c, c, device=dev, generator=g)
        s = (s + s.t()) * 0.5
        gw = torch.tensor([1e5], device=dev)
        out = torch.randn(hw, c, device=dev, generator=g)
        base = out.clone()
        ops.style_bwd(xf, s, gw, out, accumulate=True)
        ref = base + 1e5 * (xf.double() @ s.double()).float()
        ok &= report(f"style_bwd C={c} accumulate", rel_l2(out, ref), 2e-3)
        ops.style_bwd(xf, s, gw, out, accumulate=False)
        ok &= report(f"style_bwd C={c}", rel_l2(out, 1e5 * (xf.double() @ s.double()).float()), 2e-3)
    # conv + fused 2x2 max pool (epilogue shuffles) vs conv followed by the pool kernel: identical
    # bits, pre / post outputs untouched; even, odd and tiny sizes, every tile family incl. pairs
    for (hh, ww, ci, co) in [(40, 56, 64, 64), (37, 53, 64, 128), (70, 94, 128, 128), (33, 47, 128, 256),
                             (24, 40, 256, 256), (19, 8, 256, 512), (16, 16, 512, 512), (9, 11, 64, 64),
                             (64, 24, 64, 64)]:
        xx = torch.randn(hh, ww, ci, device=dev, generator=g)
        wt2 = torch.randn(co, ci, 3, 3, device=dev, generator=g) * 0.05
        wf2, _ = ops.pack_conv_weights(wt2)
        b2 = torch.randn(co, device=dev, generator=g)
        for pair_mode in (-1, 0, 1):
            ops.conv_set_tuning(pair_mode)
            pre_a = torch.full((hh, ww, co), float("nan"), device=dev)
            post_a = torch.full((hh, ww, co), float("nan"), device=dev)
            pool_a = torch.full((hh // 2, ww // 2, co), float("nan"), device=dev)
            ops.conv3x3_fwd(xx, wf2, b2, pre_a, post_a, out_pool=pool_a)
            pre_b = torch.empty_like(pre_a)
            post_b = torch.empty_like(post_a)
            pool_b = torch.empty_like(pool_a)
            ops.conv3x3_fwd(xx, wf2, b2, pre_b, post_b)
            ops.maxpool2_fwd(post_b, pool_b)
            same = torch.equal(pre_a, pre_b) and torch.equal(post_a, post_b) and torch.equal(pool_a, pool_b)
            ref_pool = F.max_pool2d(post_b.permute(2, 0, 1)[None], 2)[0].permute(1, 2, 0)
            same = same and torch.equal(pool_a, ref_pool)
            ok &= report(f"conv+fused pool {hh}x{ww} {ci}->{co} pair_mode={pair_mode}",
                         0.0 if same else 1.0, 0.0)
    ops.conv_set_tuning()
    torch.cuda.synchronize()
    return ok


def case_conv_compact() -> bool:
    """Backward bookkeeping as bits: ReLU sign bits / pool codes recorded by the forward epilogues,
    consumed by the dgrad epilogues.  Everything here must be BIT-IDENTICAL to the fp32-re-reading
    path (same accumulators, same arithmetic; only where the gate comes from differs)."""
    import torch

    from style_transfer_visualizer_b200 import ops

    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(9)
    ok = True

    def unpack_bits(bits: torch.Tensor, c: int) -> torch.Tensor:
        """[H, W, C/32] int32 -> [H, W, C] bool."""
        sh = torch.arange(32, device=bits.device, dtype=torch.int32)
        return ((bits.unsqueeze(-1) >> sh) & 1).reshape(*bits.shape[:2], c).bool()

    def route_reference(post: torch.Tensor) -> torch.Tensor:
        """[H, W, C] bool: pixel receives the pooled gradient (first maximum of its 2x2 window in
        ATen's scan order, maximum > 0); rows / columns dropped by floor mode route nothing."""
        hh, ww, c = post.shape
        ho, wo = hh // 2, ww // 2
        win = post[:2 * ho, :2 * wo].reshape(ho, 2, wo, 2, c).permute(0, 2, 4, 1, 3) \
            .reshape(ho, wo, c, 4)                               # window order a, b, d, e
        k = torch.zeros(ho, wo, c, device=post.device, dtype=torch.int64)
        m = win[..., 0].clone()
        for q in (1, 2, 3):
            better = win[..., q] > m
            k = torch.where(better, torch.full_like(k, q), k)
            m = torch.where(better, win[..., q], m)
        hit = torch.nn.functional.one_hot(k, 4).bool() & (m > 0).unsqueeze(-1)   # [ho, wo, c, 4]
        out = torch.zeros(hh, ww, c, device=post.device, dtype=torch.bool)
        out[:2 * ho, :2 * wo] = hit.reshape(ho, wo, c, 2, 2).permute(0, 3, 1, 4, 2) \
            .reshape(2 * ho, 2 * wo, c)
        return out

    def pack_bits(mask: torch.Tensor) -> torch.Tensor:
        """[H, W, C] bool -> [H, W, C/32] int32 words."""
        hh, ww, c = mask.shape
        sh = torch.arange(32, device=mask.device, dtype=torch.int64)
        packed = (mask.reshape(hh, ww, c // 32, 32).long() << sh).sum(-1)
        return torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).int().contiguous()

    # first layer: bits of conv1_1's post
    for (hh, ww) in [(64, 64), (33, 47), (70, 94)]:
        img = torch.randn(1, 3, hh, ww, device=dev, generator=g)
        w1 = torch.randn(64, 3, 3, 3, device=dev, generator=g) * 0.2
        b1 = torch.randn(64, device=dev, generator=g) * 0.1
        pre = torch.empty(hh, ww, 64, device=dev)
        post = torch.empty(hh, ww, 64, device=dev)
        pre2 = torch.empty_like(pre)
        post2 = torch.empty_like(post)
        bits = torch.full((hh, ww, 2), -1, device=dev, dtype=torch.int32)
        ops.conv3x3_first_fwd(img, w1, b1, pre, post)
        ops.conv3x3_first_fwd(img, w1, b1, pre2, post2, out_bits=bits)
        same = torch.equal(pre, pre2) and torch.equal(post, post2) and \
            torch.equal(unpack_bits(bits, 64), post > 0)
        ok &= report(f"first fwd + sign bits {hh}x{ww}", 0.0 if same else 1.0, 0.0)
        # tensor-core variant (TF32 operands): close to the exact-fp32 kernel, self-consistent bits
        pre3 = torch.full_like(pre, float("nan"))
        post3 = torch.full_like(post, float("nan"))
        bits3 = torch.full((hh, ww, 2), -1, device=dev, dtype=torch.int32)
        ops.conv3x3_first_fwd_tc(img, w1, b1, pre3, post3, out_bits=bits3)
        ok &= report(f"first fwd tensor-core pre  {hh}x{ww}", rel_l2(pre3, pre), 1.5e-3)
        ok &= report(f"first fwd tensor-core post {hh}x{ww}", rel_l2(post3, post), 1.5e-3)
        same = torch.equal(unpack_bits(bits3, 64), post3 > 0) and \
            torch.equal(post3, torch.relu(post3)) and \
            int((post3.view(torch.int32) & 0x1FFF).abs().max()) == 0
        ok &= report(f"first fwd tensor-core bits / rounding {hh}x{ww}", 0.0 if same else 1.0, 0.0)
        # haloed band: rows 1..hh-2 of the same image as a band with one halo row above and below
        band = img[:, :, 0:hh, :].contiguous()
        pre4 = torch.full((hh - 2, ww, 64), float("nan"), device=dev)
        ops.conv3x3_first_fwd_tc(band, w1, b1, pre4, None, rows=hh - 2, in_row0=1)
        ok &= report(f"first fwd tensor-core band rows {hh}x{ww}",
                     0.0 if torch.equal(pre4, pre3[1:hh - 1]) else 1.0, 0.0)

    for (hh, ww, ci, co) in [(40, 56, 64, 64), (37, 53, 64, 128), (70, 94, 128, 128), (33, 47, 128, 256),
                             (24, 40, 256, 256), (19, 8, 256, 512), (16, 16, 512, 512), (9, 11, 64, 64),
                             (135, 240, 64, 64)]:
        xx = torch.randn(hh, ww, ci, device=dev, generator=g)
        wt = torch.randn(co, ci, 3, 3, device=dev, generator=g) * 0.05
        wf, wd = ops.pack_conv_weights(wt)
        bb = torch.randn(co, device=dev, generator=g) * 0.3
        for pair_mode in (-1, 0, 1):
            ops.conv_set_tuning(pair_mode)
            tag = f"{hh}x{ww} {ci}->{co} pair_mode={pair_mode}"
            # ---- forward: sign bits
            post = torch.empty(hh, ww, co, device=dev)
            post2 = torch.empty_like(post)
            bits = torch.full((hh, ww, co // 32), -1, device=dev, dtype=torch.int32)
            ops.conv3x3_fwd(xx, wf, bb, None, post)
            ops.conv3x3_fwd(xx, wf, bb, None, post2, out_bits=bits)
            same = torch.equal(post, post2) and torch.equal(unpack_bits(bits, co), post > 0)
            ok &= report(f"fwd + sign bits {tag}", 0.0 if same else 1.0, 0.0)
            # ---- forward: pool + routing bits, post not stored
            ho, wo = hh // 2, ww // 2
            pool = torch.full((ho, wo, co), float("nan"), device=dev)
            code = torch.full((hh, ww, co // 32), -1, device=dev, dtype=torch.int32)
            ops.conv3x3_fwd(xx, wf, bb, None, None, out_pool=pool, out_code=code)
            pool_ref = torch.empty_like(pool)
            ops.maxpool2_fwd(post, pool_ref)
            same = torch.equal(pool, pool_ref) and torch.equal(unpack_bits(code, co),
                                                               route_reference(post))
            ok &= report(f"fwd + pool + route bits {tag}", 0.0 if same else 1.0, 0.0)
            # ---- dgrad gated by bits == dgrad gated by the fp32 activation (layer gated: [hh,ww,ci])
            dy = torch.randn(hh, ww, co, device=dev, generator=g)
            act = torch.randn(hh, ww, ci, device=dev, generator=g).relu()
            abits = pack_bits(act > 0)
            if ci % 64 == 0:
                prev = torch.randn(hh, ww, ci, device=dev, generator=g)
                for accumulate in (False, True):
                    d_a = prev.clone()
                    d_b = prev.clone()
                    ops.conv3x3_dgrad(dy, wd, d_a, relu_src=act, accumulate=accumulate)
                    ops.conv3x3_dgrad(dy, wd, d_b, relu_bits=abits, accumulate=accumulate)
                    ok &= report(f"dgrad gate bits acc={int(accumulate)} {tag}",
                                 0.0 if torch.equal(d_a, d_b) else 1.0, 0.0)
        ops.conv_set_tuning()
    # ---- dgrad fused with pool + ReLU backward == dgrad -> maxpool2_bwd(relu_mask); odd sizes: the
    # dropped last row / column stays zero
    for (h2, w2, cpool, cnext) in [(40, 56, 64, 128), (37, 53, 64, 128), (71, 95, 128, 256),
                                   (33, 47, 256, 512), (16, 16, 512, 512), (135, 241, 64, 128)]:
        ho, wo = h2 // 2, w2 // 2
        post = torch.randn(h2, w2, cpool, device=dev, generator=g).relu()
        post = torch.where(torch.rand(h2, w2, cpool, device=dev, generator=g) < 0.1,
                           post.roll(1, 1), post)            # some exact ties inside windows
        wt = torch.randn(cnext, cpool, 3, 3, device=dev, generator=g) * 0.05
        _wf, wd = ops.pack_conv_weights(wt)
        dy = torch.randn(ho, wo, cnext, device=dev, generator=g)
        code = pack_bits(route_reference(post))
        for pair_mode in (-1, 0, 1):
            ops.conv_set_tuning(pair_mode)
            d_pool = torch.empty(ho, wo, cpool, device=dev)
            ops.conv3x3_dgrad(dy, wd, d_pool)
            want = torch.empty(h2, w2, cpool, device=dev)
            ops.maxpool2_bwd(d_pool, post, want, relu_mask=True)
            got = torch.zeros(h2, w2, cpool, device=dev)
            ops.conv3x3_dgrad_unpool(dy, wd, code, got)
            ok &= report(f"dgrad + unpool {h2}x{w2} {cnext}->{cpool} pair_mode={pair_mode}",
                         0.0 if torch.equal(got, want) else 1.0, 0.0)
        ops.conv_set_tuning()
    # ---- dgrad + ReLU gate + Gram backward of the tapped layer in one launch (second accumulator)
    # == style_bwd followed by the accumulating dgrad, up to fp32 rounding of the final addition
    for (hh, ww, cn, cd) in [(64, 96, 64, 64), (135, 240, 64, 64), (70, 94, 128, 128),
                             (200, 304, 128, 128), (24, 40, 128, 256), (33, 47, 64, 128)]:
        dy = torch.randn(hh, ww, cd, device=dev, generator=g)
        wt = torch.randn(cd, cn, 3, 3, device=dev, generator=g) * 0.05
        _wf, wd = ops.pack_conv_weights(wt)
        feat = torch.randn(hh, ww, cn, device=dev, generator=g)
        sm = torch.randn(cn, cn, device=dev, generator=g) * 0.1
        sm = (sm + sm.t()) * 0.5
        gw = torch.tensor([3.0], device=dev)
        abits = pack_bits(torch.rand(hh, ww, cn, device=dev, generator=g) > 0.4)
        want = torch.empty(hh, ww, cn, device=dev)
        ops.style_bwd(feat, sm, gw, want, accumulate=False)
        ops.conv3x3_dgrad(dy, wd, want, relu_bits=abits, accumulate=True)
        got = torch.full((hh, ww, cn), float("nan"), device=dev)
        ops.conv3x3_dgrad_style(dy, wd, got, relu_bits=abits, feat=feat, s_mat=sm, grad_w=gw)
        ok &= report(f"dgrad + gate + fused style bwd {hh}x{ww} {cd}->{cn}", rel_l2(got, want), 1e-3)
    torch.cuda.synchronize()
    return ok


def case_halo() -> bool:
    """stv_halo_exchange on ONE GPU: the neighbours' buffers and flag words are ordinary device
    memory of this process and their side of the hand-shake is played by the host (their flags are
    raised before the launch), so the data path -- boundary rows pushed into the neighbours' halo
    rows, zero padding at the image edge, multi-plane NCHW rows, unequal band sizes -- and the
    flags this rank raises are checked exactly.  The concurrent hand-shake itself needs two GPUs
    (tools/sharded_check.py): two mutually waiting kernels of one GPU can be serialised behind each
    other by a shared hardware queue."""
    import torch

    from style_transfer_visualizer_b200 import ops

    dev = torch.device("cuda")
    ok = True
    slot = 5
    for planes, rows_of, cols, ch in [(1, (8, 8, 6), 20, 64), (1, (5, 7, 3), 12, 128),
                                      (3, (16, 16, 12), 64, 1), (1, (4, 4, 4), 960, 64)]:
        n = len(rows_of)
        row_floats = cols * ch
        flags = [torch.zeros(64 * 4, device=dev, dtype=torch.int32) for _ in range(n)]
        epoch = [torch.zeros(64, device=dev, dtype=torch.int32) for _ in range(n)]
        done = [torch.zeros(64, device=dev, dtype=torch.int32) for _ in range(n)]
        for it in range(1, 5):      # four epochs; the last two without the "ready" round trip
            for r in range(n):
                bufs = [torch.randn(planes, rows + 2, row_floats, device=dev) for rows in rows_of]
                before = [b.clone() for b in bufs]
                flags[r][4 * slot:4 * slot + 4] = it        # the neighbours "have signalled"
                torch.cuda.synchronize()
                ops.halo_exchange(
                    bufs[r], up_ptr=bufs[r - 1].data_ptr() if r > 0 else None,
                    down_ptr=bufs[r + 1].data_ptr() if r < n - 1 else None,
                    rows=rows_of[r], rows_up=rows_of[r - 1] if r > 0 else 0,
                    rows_down=rows_of[r + 1] if r < n - 1 else 0, row_floats=row_floats,
                    planes=planes, flags_mine=flags[r],
                    flags_up_ptr=flags[r - 1].data_ptr() if r > 0 else None,
                    flags_down_ptr=flags[r + 1].data_ptr() if r < n - 1 else None,
                    epoch=epoch[r], done=done[r], slot=slot, wait_ready=it <= 2)
                torch.cuda.synchronize()
                want = [b.clone() for b in before]
                if r > 0:
                    want[r - 1][:, rows_of[r - 1] + 1] = before[r][:, 1]
                else:
                    want[r][:, 0] = 0.0
                if r < n - 1:
                    want[r + 1][:, 0] = before[r][:, rows_of[r]]
                else:
                    want[r][:, rows_of[r] + 1] = 0.0
                same = all(torch.equal(bufs[k], want[k]) for k in range(n))
                same = same and int(epoch[r][slot]) == it and int(done[r][slot]) == 0
                if r > 0:       # B (and A when requested) "from below" raised at the rank above
                    same = same and flags[r - 1][4 * slot + 3].item() == it \
                        and (it > 2 or flags[r - 1][4 * slot + 1].item() == it)
                if r < n - 1:   # B (and A) "from above" raised at the rank below
                    same = same and flags[r + 1][4 * slot + 2].item() == it \
                        and (it > 2 or flags[r + 1][4 * slot + 0].item() == it)
                ok &= report(f"halo push planes={planes} rows={rows_of} row={row_floats} rank={r} "
                             f"epoch={it}", 0.0 if same else 1.0, 0.0)
    return ok


def case_gram() -> bool:
    import torch

    from style_transfer_visualizer_b200 import ops

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    ok = True
    for (hw, c) in [(64 * 64, 64), (33 * 47, 64), (1, 64), (32 * 32, 128), (33 * 47, 128),
                    (135 * 240, 256), (67 * 120, 512), (7, 512), (540 * 960, 128)]:
        x = torch.randn(hw, c, device=dev, generator=g) + 0.3
        ws = ops.gram_workspace(hw, c, dev)
        gram = torch.full((c, c), float("nan"), device=dev)
        ops.gram_loss_fwd(x, ws, gram_out=gram)
        f64 = x.double().t()
        r = f64 @ f64.t()
        ref = r.clamp(max=5e5) / (c * hw)
        good = report(f"gram hw={hw} C={c}", rel_l2(gram, ref), 1e-3)
        if not good:
            print("   got", gram[0, :4].tolist(), gram[c - 1, c - 4:].tolist())
            print("   ref", ref[0, :4].tolist(), ref[c - 1, c - 4:].tolist())
        ok &= good
        # loss + S against autograd of the reference formula
        target = (ref * 0.9 + 0.01).float()
        loss = torch.zeros(1, device=dev)
        s = torch.full((c, c), float("nan"), device=dev)
        ops.gram_loss_fwd(x, ws, target=target, s_out=s, loss_out=loss)
        xa = x.double().requires_grad_(True)
        fa = xa.t()
        ga = (fa @ fa.t()).clamp(max=5e5) / (c * hw)
        la = torch.nn.functional.mse_loss(ga, target.double())
        la.backward()
        ok &= report(f"gram loss hw={hw} C={c}", abs(float(loss) - float(la.detach())) / float(la.detach()), 2e-2)
        # dX = X @ S
        ok &= report(f"gram S -> dX hw={hw} C={c}", rel_l2(x.double() @ s.double(), xa.grad), 2e-2)
    # clamp-active case
    hw, c = 4096, 128
    x = torch.randn(hw, c, device=dev, generator=g) * 12 + 3
    ws = ops.gram_workspace(hw, c, dev)
    gram = torch.empty(c, c, device=dev)
    s = torch.empty(c, c, device=dev)
    loss = torch.zeros(1, device=dev)
    target = torch.rand(c, c, device=dev, generator=g)
    ops.gram_loss_fwd(x, ws, target=target, gram_out=gram, s_out=s, loss_out=loss)
    r = x.double().t() @ x.double()
    frac = float((r > 5e5).double().mean())
    # entries within TF32 error of the clamp threshold may legitimately land on either side
    safe = ((r - 5e5).abs() > 5e5 * 5e-3)
    ref = r.clamp(max=5e5) / (c * hw)
    ok &= report(f"gram clamp-active ({frac:.2f} clamped)", rel_l2(gram, ref), 1e-3)
    sref = 4.0 / (c * c * c * hw) * (r <= 5e5) * (ref - target.double())
    ok &= report("gram S clamp mask", rel_l2(s.double() * safe, sref * safe), 2e-2)
    torch.cuda.synchronize()
    return ok


def _time(fn, iters=10, warm=3):  # noqa: ANN001, ANN202
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case_perf() -> bool:
    import torch
    import torch.nn.functional as F  # noqa: N812

    from style_transfer_visualizer_b200 import ops

    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(4)
    layers = [  # 1080p VGG19 conv shapes (H, W, Cin, Cout)
        (1080, 1920, 64, 64), (540, 960, 64, 128), (540, 960, 128, 128), (270, 480, 128, 256),
        (270, 480, 256, 256), (135, 240, 256, 512), (135, 240, 512, 512), (67, 120, 512, 512),
        (512, 512, 64, 64), (256, 256, 128, 128), (128, 128, 256, 256), (64, 64, 512, 512),
        (32, 32, 512, 512),
    ]
    torch.backends.cudnn.allow_tf32 = True
    for (h, w, cin, cout) in layers:
        x = torch.randn(h, w, cin, device=dev, generator=g)
        wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        b = torch.zeros(cout, device=dev)
        wf, wd = ops.pack_conv_weights(wt)
        post = torch.empty(h, w, cout, device=dev)
        flops = 2.0 * 9 * cin * cout * h * w
        ms = _time(lambda: ops.conv3x3_fwd(x, wf, b, None, post))  # noqa: B023
        line = f"PERF conv {h}x{w} {cin}->{cout}: auto {ms:.3f} ms {flops / ms / 1e9:.1f} TF/s"
        for bn in (64, 128, 256):
            if cout % bn:
                continue
            for mh in (1, 2):
                for tw in (8, 16, 32):
                    try:
                        ms2 = _time(lambda: ops.conv_igemm2_ex(x, wf, taps=9, bias=b, out_post=post, block_n=bn, m_halves=mh, tw=tw))  # noqa: B023
                        line += f" | v2 n{bn}m{mh}w{tw} {flops / ms2 / 1e9:.0f}"
                    except Exception:  # noqa: BLE001
                        line += f" | v2 n{bn}m{mh}w{tw} n/a"
        if cin % 64 == 0:
            dyb = torch.randn(h, w, cout, device=dev, generator=g)
            dxb = torch.zeros(h, w, cin, device=dev)
            act = torch.randn(h, w, cin, device=dev, generator=g)
            ms4 = _time(lambda: ops.conv3x3_dgrad(dyb, wd, dxb, relu_src=act, accumulate=True))  # noqa: B023
            line += f" | dgrad+gate+acc auto {ms4:.3f} ms {flops / ms4 / 1e9:.0f}"
        xt = x.permute(2, 0, 1).unsqueeze(0).contiguous(memory_format=torch.channels_last)
        wcl = wt.contiguous(memory_format=torch.channels_last)
        ms3 = _time(lambda: F.relu(F.conv2d(xt, wcl, b, padding=1)))  # noqa: B023
        line += f" | cudnn-tf32 {ms3:.3f} ms {flops / ms3 / 1e9:.1f}"
        print(line, flush=True)
    for (hw, c) in [(1080 * 1920, 64), (540 * 960, 128), (270 * 480, 256), (135 * 240, 512),
                    (67 * 120, 512), (512 * 512, 64), (64 * 64, 512)]:
        x = torch.randn(hw, c, device=dev, generator=g)
        ws = ops.gram_workspace(hw, c, dev)
        tgt = torch.zeros(c, c, device=dev)
        s = torch.empty(c, c, device=dev)
        loss = torch.zeros(1, device=dev)
        ms = _time(lambda: ops.gram_loss_fwd(x, ws, target=tgt, s_out=s, loss_out=loss))  # noqa: B023
        gb = hw * c * 4 / 1e9
        xt = x.t().contiguous()
        ms2 = _time(lambda: torch.mm(xt, xt.t()))  # noqa: B023
        print(f"PERF gram hw={hw} C={c}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  "
              f"{2.0 * c * c * hw / ms / 1e9:.1f} TF/s(full) | torch.mm fp32 {ms2:.3f} ms", flush=True)
        gw = torch.ones(1, device=dev)
        dy = torch.empty(hw, c, device=dev)
        ms3 = _time(lambda: ops.style_bwd(x, s, gw, dy, accumulate=False))  # noqa: B023
        print(f"PERF style_bwd hw={hw} C={c}: {ms3:.3f} ms {2.0 * c * c * hw / ms3 / 1e9:.1f} TF/s", flush=True)
    # memory-bound kernels at 1080p
    h, w, c = 1080, 1920, 64
    x = torch.randn(h, w, c, device=dev, generator=g).relu()
    y = torch.empty(h // 2, w // 2, c, device=dev)
    ms = _time(lambda: ops.maxpool2_fwd(x, y))
    print(f"PERF maxpool fwd 1080p C=64: {ms:.3f} ms {(x.numel() + y.numel()) * 4 / ms / 1e6:.0f} GB/s")
    dx = torch.empty_like(x)
    ms = _time(lambda: ops.maxpool2_bwd(y, x, dx, relu_mask=True))
    print(f"PERF maxpool bwd 1080p C=64: {ms:.3f} ms {(2 * x.numel() + y.numel()) * 4 / ms / 1e6:.0f} GB/s")
    img = torch.randn(1, 3, h, w, device=dev, generator=g)
    w1 = torch.randn(64, 3, 3, 3, device=dev, generator=g)
    b1 = torch.zeros(64, device=dev)
    pre = torch.empty(h, w, 64, device=dev)
    post = torch.empty(h, w, 64, device=dev)
    ms = _time(lambda: ops.conv3x3_first_fwd(img, w1, b1, pre, post))
    print(f"PERF conv_first_fwd 1080p: {ms:.3f} ms {(2 * pre.numel() + img.numel()) * 4 / ms / 1e6:.0f} GB/s")
    dimg = torch.empty_like(img)
    ms = _time(lambda: ops.conv3x3_first_dgrad(pre, w1, dimg))
    print(f"PERF conv_first_dgrad 1080p: {ms:.3f} ms {(pre.numel() + img.numel()) * 4 / ms / 1e6:.0f} GB/s")
    w16 = ops.pack_first_dgrad_weights(w1)
    ms = _time(lambda: ops.conv3x3_first_dgrad_tc(pre, w16, dimg))
    print(f"PERF conv_first_dgrad_tc 1080p: {ms:.3f} ms {(pre.numel() + img.numel()) * 4 / ms / 1e6:.0f} GB/s")
    wrows = ops.pack_first_dgrad_rows(w1)
    ms = _time(lambda: ops.conv3x3_first_dgrad_rows(pre, wrows, dimg))
    print(f"PERF conv_first_dgrad_rows 1080p: {ms:.3f} ms {(pre.numel() + img.numel()) * 4 / ms / 1e6:.0f} GB/s")
    p = img.flatten().clone()
    gr = torch.randn_like(p)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    ms = _time(lambda: ops.adam_step(p, gr, m, v, beta1=0.9, beta2=0.999, eps=1e-8, step_size=0.01, bias2_sqrt=1.0))
    print(f"PERF adam 1080p: {ms:.3f} ms {p.numel() * 28 / ms / 1e6:.0f} GB/s")
    out8 = torch.empty(h, w, 3, device=dev, dtype=torch.uint8)
    ms = _time(lambda: ops.frame_to_u8(img, out8, denormalize=True))
    print(f"PERF frame_to_u8 1080p: {ms:.3f} ms {(img.numel() * 4 + out8.numel()) / ms / 1e6:.0f} GB/s")
    return True


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--case", default=None, help="internal: run one case in-process")
    ap.add_argument("--out", default=None)
    ap.add_argument("--timeout", type=int, default=300)
    args = ap.parse_args()
    if args.case:
        ok = globals()[f"case_{args.case}"]()
        print(f"CASE {args.case}: {'OK' if ok else 'FAILED'}", flush=True)
        return 0 if ok else 1
    rc_all = 0
    out_f = open(args.out, "w") if args.out else None  # noqa: SIM115
    for case in args.cases.split(","):
        t0 = time.time()
        try:
            proc = subprocess.run([sys.executable, __file__, "--case", case], capture_output=True,
                                  text=True, timeout=args.timeout, check=False)
            text = proc.stdout + proc.stderr[-4000:]
            rc = proc.returncode
        except subprocess.TimeoutExpired as exc:
            text = (exc.stdout or b"").decode() if isinstance(exc.stdout, bytes) else (exc.stdout or "")
            text += f"\nCASE {case}: TIMEOUT after {args.timeout}s\n"
            rc = 124
        text += f"[case {case} rc={rc} {time.time() - t0:.1f}s]\n"
        print(text, flush=True)
        if out_f:
            out_f.write(text)
            out_f.flush()
        rc_all |= rc
    return 1 if rc_all else 0


if __name__ == "__main__":
    raise SystemExit(main())
