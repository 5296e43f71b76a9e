"""Kernel-level parity on a real B200: every C-ABI entry point against plain PyTorch fp32.

The cases live in tools/gpu_selftest.py (also usable stand-alone under gpurun); each runs in its
own subprocess so a trapping kernel cannot poison the remaining tests' CUDA context.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["elementwise", "conv", "conv2", "conv_variants", "conv_compact", "halo", "gram"])
def test_kernel_case(case: str) -> None:
    proc = subprocess.run([sys.executable, str(ROOT / "tools" / "gpu_selftest.py"), "--case", case],
                          capture_output=True, text=True, timeout=600, check=False, cwd=ROOT)
    fails = [ln for ln in proc.stdout.splitlines() if ln.startswith("FAIL")]
    assert proc.returncode == 0 and not fails, \
        "\n".join(fails) + "\n" + proc.stdout[-3000:] + proc.stderr[-3000:]
    assert proc.stdout.count("PASS") > 5


@pytest.mark.gpu
def test_conv_cta_pair_matches_single_cta_bit_for_bit(cuda_device) -> None:  # noqa: ANN001
    """The cta_group::2 variant (M = 256 per instruction, weight tile split over two CTAs' shared
    memory) accumulates every output in the same K order as the single-CTA kernel: identical bits,
    for every tile shape, ring depth and taps-per-stage setting, ragged sizes and an odd patch
    count (last pair half empty) included; and both agree with the CUDA-core cross-check."""
    import torch

    from style_transfer_visualizer_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(11)
    ops.conv_set_split(0)  # the split-K second issuer sums in a different order (tested below)
    try:
        for h, w, c, n in [(33, 47, 128, 256), (24, 40, 256, 128), (19, 21, 64, 64), (16, 16, 512, 512)]:
            x = torch.randn(h, w, c, device=cuda_device, generator=g)
            wt = torch.randn(n, c, 3, 3, device=cuda_device, generator=g) * 0.05
            wf, _ = ops.pack_conv_weights(wt)
            bias = torch.randn(n, device=cuda_device, generator=g)
            mask = torch.randn(h, w, n, device=cuda_device, generator=g)
            add = torch.randn(h, w, n, device=cuda_device, generator=g)
            ref = ops.conv_ref(x, wf, bias, taps=9, relu=False)
            for bn in (256, 128, 64):
                if n % bn:
                    continue
                for mh in (1, 2):
                    outs = []
                    for pair, depth, tps in [(0, 0, 0), (1, 0, 0), (1, 2, 1), (1, 3, 3), (0, 3, 1)]:
                        ops.conv_set_tuning(pair, depth, depth, tps)
                        pre = torch.full((h, w, n), float("nan"), device=cuda_device)
                        post = torch.full((h, w, n), float("nan"), device=cuda_device)
                        ops.conv_igemm2_ex(x, wf, taps=9, bias=bias, out_pre=pre, out_post=post,
                                           block_n=bn, m_halves=mh)
                        heavy = torch.full((h, w, n), float("nan"), device=cuda_device)
                        ops.conv_igemm2_ex(x, wf, taps=9, mask_src=mask, add_src=add, out_pre=heavy,
                                           block_n=bn, m_halves=mh)
                        outs.append((pre, post, heavy))
                    for o in outs[1:]:
                        for a, b in zip(outs[0], o):
                            assert torch.equal(a, b)
                    assert torch.allclose(outs[0][0], ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
                    assert torch.equal(outs[0][1], torch.relu(outs[0][0]))
    finally:
        ops.conv_set_tuning()
        ops.conv_set_split()


@pytest.mark.gpu
def test_conv_split_k_second_issuer_matches_single_issuer(cuda_device) -> None:  # noqa: ANN001
    """One-half tiles with the split-K second issuer (two partial accumulators added in the
    epilogue) against the single-issuer kernel: same products, fp32 summation order differs ->
    agreement to fp32 round-off of a K = 9 * C term sum (2e-5 of the largest output); every epilogue (plain, gate bits,
    accumulate, 1x1) and ring depth, an odd number of ring stages per tile included."""
    import torch

    from style_transfer_visualizer_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(12)
    try:
        for h, w, c, n, taps in [(24, 40, 256, 128, 9), (19, 21, 64, 64, 9), (16, 16, 512, 512, 9),
                                 (33, 47, 96, 128, 9), (20, 20, 64, 64, 1), (17, 31, 32, 64, 1),
                                 (64, 64, 512, 512, 9),
                                 # more tiles than CTAs: several tiles per CTA, even and odd stage counts
                                 (128, 160, 64, 128, 9), (128, 160, 96, 64, 9), (128, 320, 96, 128, 9)]:
            x = torch.randn(h, w, c, device=cuda_device, generator=g)
            if taps == 9:
                wt = torch.randn(n, c, 3, 3, device=cuda_device, generator=g) * 0.05
                wf, _ = ops.pack_conv_weights(wt)
            else:
                wf = (torch.randn(1, n, c, device=cuda_device, generator=g) * 0.05).contiguous()
            bias = torch.randn(n, device=cuda_device, generator=g)
            mask = torch.randn(h, w, n, device=cuda_device, generator=g)
            add = torch.randn(h, w, n, device=cuda_device, generator=g)
            for bn in (128, 64):
                if n % bn:
                    continue
                for depth in (0, 2, 3):
                    outs = []
                    for split in (0, 1):
                        ops.conv_set_split(split)
                        ops.conv_set_tuning(0, depth, depth, 0)
                        pre = torch.full((h, w, n), float("nan"), device=cuda_device)
                        post = torch.full((h, w, n), float("nan"), device=cuda_device)
                        ops.conv_igemm2_ex(x, wf, taps=taps, bias=bias, out_pre=pre, out_post=post,
                                           block_n=bn, m_halves=1)
                        heavy = torch.full((h, w, n), float("nan"), device=cuda_device)
                        ops.conv_igemm2_ex(x, wf, taps=taps, mask_src=mask, add_src=add,
                                           out_pre=heavy, block_n=bn, m_halves=1)
                        outs.append((pre, post, heavy))
                    scale = float(outs[0][0].abs().max())
                    for a, b in zip(outs[0], outs[1]):
                        assert torch.isfinite(b).all()
                        assert float((a - b).abs().max()) <= 2e-5 * scale + 1e-6, (h, w, c, n, taps, bn, depth)
                    assert torch.equal(outs[1][1], torch.relu(outs[1][0]))
    finally:
        ops.conv_set_tuning()
        ops.conv_set_split()


@pytest.mark.gpu
def test_conv_weight_stationary_is_bit_identical(cuda_device) -> None:  # noqa: ANN001
    """64 -> 64 layers with several tiles per CTA keep their weights resident in shared memory (CTA
    pairs, ring stages carry activations only).  Same MMAs in the same order as the streaming
    kernel of the same tile family: every output -- forward with fused pool / route bits / sign
    bits, bit-gated dgrad, dgrad with the fused style backward -- must match bit for bit; and the
    forward agrees with the CUDA-core cross-check."""
    import torch

    from style_transfer_visualizer_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(13)
    try:
        for h, w in [(300, 400), (301, 403), (512, 512)]:
            x = torch.randn(h, w, 64, device=cuda_device, generator=g).relu()
            wt = torch.randn(64, 64, 3, 3, device=cuda_device, generator=g) * 0.05
            wf, wd = ops.pack_conv_weights(wt)
            bias = torch.randn(64, device=cuda_device, generator=g)
            dy = torch.randn(h, w, 64, device=cuda_device, generator=g)
            gate = ops.relu_bits_buffer(h, w, 64, cuda_device)
            gate.random_(-2 ** 31, 2 ** 31 - 1)
            feat = torch.randn(h, w, 64, device=cuda_device, generator=g)
            smat = torch.randn(64, 64, device=cuda_device, generator=g)
            smat = ((smat + smat.t()) * 0.5).contiguous()
            gl = torch.tensor([0.37], device=cuda_device)
            outs = []
            for mode in (0, -1):
                ops.conv_set_resident(mode)
                ops.conv_plan_override()
                if mode == 0:  # same tile family as the stationary rule picks: two halves on a CTA pair
                    for bwd in (False, True):
                        ops.conv_plan_override(h, w, 64, 64, backward=bwd, block_n=64, m_halves=2, pair=1)
                post = torch.full((h, w, 64), float("nan"), device=cuda_device)
                pool = torch.full((h // 2, w // 2, 64), float("nan"), device=cuda_device)
                bits = ops.relu_bits_buffer(h, w, 64, cuda_device).fill_(-1)
                code = ops.pool_code_buffer(h, w, 64, cuda_device).fill_(-1)
                ops.conv3x3_fwd(x, wf, bias, None, post, out_pool=pool, out_bits=bits, out_code=code)
                dx = torch.full((h, w, 64), float("nan"), device=cuda_device)
                ops.conv3x3_dgrad(dy, wd, dx, relu_bits=gate)
                dxs = torch.full((h, w, 64), float("nan"), device=cuda_device)
                ops.conv3x3_dgrad_style(dy, wd, dxs, relu_bits=gate, feat=feat, s_mat=smat, grad_w=gl)
                outs.append((post, pool, bits, code, dx, dxs))
            for a, b in zip(outs[0], outs[1]):
                assert torch.equal(a, b)
            ref = ops.conv_ref(x, wf, bias, taps=9, relu=True)
            assert torch.allclose(outs[1][0], ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    finally:
        ops.conv_set_resident()
        ops.conv_plan_override()
