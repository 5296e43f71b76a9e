"""Row-band sharding check (run under torchrun on >= 2 GPUs): the sharded model must reproduce the
single-GPU model's losses and input gradient, and a sharded Adam run must track the unsharded one.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_check.py
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from style_transfer_visualizer_b200 import jobs, synthetic  # noqa: E402
from style_transfer_visualizer_b200.optim import FusedAdam  # noqa: E402
from style_transfer_visualizer_b200.sharded import ShardedStyleContentModel, plan_bands  # noqa: E402


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def golden_4k(info, dev) -> bool:  # noqa: ANN001
    """The SHARDED first closure at 3840x2160 against tests/golden/adam_random_4k_c5.npz -- outputs
    of the unmodified reference on the CPU (oracle/make_golden.py), not of this package's
    single-GPU path.  Gates = 1.5 x the single-GPU deviation from the same fixture."""
    import numpy as np

    gold = np.load(ROOT / "tests" / "golden" / "adam_random_4k_c5.npz", allow_pickle=False)
    h, w = 2160, 3840
    feats = synthetic.random_vgg19_features(0)
    content = synthetic.synthetic_image(1, h, w)
    style = synthetic.synthetic_image(2, h, w)
    start = torch.randn(content.shape, generator=torch.Generator().manual_seed(3))
    model = ShardedStyleContentModel(feats, [0, 5, 10, 19, 28], [21], dev)
    model.set_targets(style, content)
    x = model.band_of(start).requires_grad_(True)
    sl, cl = model(x)
    (1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()).backward()
    full = model.gather_image(x.grad)
    if info.rank != 0:
        return True
    g = full.detach().cpu().numpy()[..., ::8, ::8].astype(np.float64)
    ref = gold["first_grad"].astype(np.float64)
    e_grad = float(np.linalg.norm(g - ref) / np.linalg.norm(ref))
    ls = np.array([float(v.detach()) for v in sl])
    lc = np.array([float(v.detach()) for v in cl])
    e_style = float(np.max(np.abs(ls - gold["layer_style"]) / gold["layer_style"]))
    e_content = float(np.max(np.abs(lc - gold["layer_content"]) / gold["layer_content"]))
    good = e_grad <= 0.034 and e_style <= 5.1e-4 and e_content <= 1e-3
    print(f"{'PASS' if good else 'FAIL'} sharded x{info.world_size} 3840x2160 vs reference golden "
          f"(halo: {model.engine.halo_mode}): grad_rel_l2={e_grad:.3e} layer_style_rel_max="
          f"{e_style:.3e} content_rel={e_content:.3e}", flush=True)
    return good


def main() -> int:
    info = jobs.init_distributed()
    dev = torch.device("cuda", info.local_rank)
    torch.cuda.set_device(dev)
    ok = True
    sizes = [(160, 96, 160, 96), (150, 112, 96, 80)]
    if "--big" in sys.argv:
        sizes = [(1080, 1920, 1080, 1920)]
    for (h, w, sh, sw) in sizes:
        if min(h, sh) < 16 * info.world_size:
            continue  # fewer 16-row units than ranks
        feats = synthetic.random_vgg19_features(0)
        content = synthetic.synthetic_image(1, h, w)
        style = synthetic.synthetic_image(2, sh, sw)
        start = content + 0.1 * torch.randn(content.shape, generator=torch.Generator().manual_seed(3))
        model = ShardedStyleContentModel(feats, [0, 5, 10, 19, 28], [21], dev)
        model.set_targets(style, content)
        x = model.band_of(start).requires_grad_(True)
        sl, cl = model(x)
        loss = 1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()
        loss.backward()
        full_grad = model.gather_image(x.grad)
        losses = torch.stack([v.detach() for v in sl + cl])
        # sharded Adam: 5 steps
        xs = model.band_of(start).requires_grad_(True)
        opt = FusedAdam([xs], lr=0.01)
        hist = []

        def closure(xs=xs, opt=opt, model=model, hist=hist):
            opt.zero_grad()
            s_, c_ = model(xs)
            total = 1e5 * torch.stack(s_).sum() + torch.stack(c_).sum()
            total.backward()
            hist.append(float(total.detach()))
            return total

        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            opt.step(closure)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        final = model.gather_image(xs)
        if info.rank == 0:
            import style_transfer_visualizer_b200.core_model as cm

            original = cm.initialize_vgg
            cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
            try:
                ref = cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(dev)
            finally:
                cm.initialize_vgg = original
            ref.set_targets(style.to(dev), content.to(dev))
            xr = start.to(dev).requires_grad_(True)
            rs, rc = ref(xr)
            (1e5 * torch.stack(rs).sum() + torch.stack(rc).sum()).backward()
            rl = torch.stack([v.detach() for v in rs + rc])
            e_loss = float(((losses - rl).abs() / (rl.abs() + 1e-30)).max())
            e_grad = rel(full_grad, xr.grad)
            xa = start.to(dev).requires_grad_(True)
            oa = FusedAdam([xa], lr=0.01)
            rh = []

            def rclosure():
                oa.zero_grad()
                s_, c_ = ref(xa)
                total = 1e5 * torch.stack(s_).sum() + torch.stack(c_).sum()
                total.backward()
                rh.append(float(total.detach()))
                return total

            for _ in range(5):
                oa.step(rclosure)
            e_hist = max(abs(a - b) / abs(b) for a, b in zip(hist, rh))
            e_final = rel(final, xa.detach())
            good = e_loss < 1e-4 and e_grad < 1e-3 and e_hist < 1e-3 and e_final < 1e-3
            ok &= good
            print(f"{'PASS' if good else 'FAIL'} sharded x{info.world_size} {h}x{w} (style {sh}x{sw}) "
                  f"bands={plan_bands(h, info.world_size)} loss_rel={e_loss:.2e} grad_rel={e_grad:.2e} "
                  f"adam_loss_rel={e_hist:.2e} final_rel={e_final:.2e} step={dt * 1e3:.2f} ms",
                  flush=True)
    if "--golden4k" in sys.argv:
        ok &= golden_4k(info, dev)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    jobs.barrier()
    jobs.shutdown()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    raise SystemExit(main())
