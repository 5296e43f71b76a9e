"""End-to-end parity of the CUDA path with the reference, on a real B200.

Each golden fixture (tests/golden/*.npz) holds what the UNMODIFIED reference produced on the CPU in
fp32 for a seeded case: per-layer losses and input gradient of the first closure, the loss
history of ``OptimizationRunner.run()``, the final image and the timelapse frames.  The same case
is run here through this package's ``StyleContentModel`` + ``OptimizationRunner`` + fused
optimisers (which call the C-ABI kernels), eagerly and through the whole-step CUDA graph.

Stated tolerance (TF32 multiply / FP32 accumulate in the convolutions, the Gram contraction and the
style backward; everything else fp32) -- the reference itself runs its convolutions in TF32 on
CUDA (torch.backends.cudnn.allow_tf32 defaults to True).  The kernels are bit-reproducible, so
every fixture has its OWN gate: 1.5 x the deviation measured on a B200 (profiles/
r2_parity_report_v1.log), never looser than the global ceilings below.  A regression that doubles
an error fails.
    per-layer style loss      relative, ceiling 2e-3   (measured 1.4e-4 .. 1.2e-3)
    per-layer content loss    absolute 1e-4 x total loss   (exactly 0 for content init)
    input gradient            relative L2, ceiling 3.5e-2 (measured 3.8e-3 .. 1.5e-2 up to 1080p,
                              2.2e-2 at 3840x2160), cosine >= 0.9995.  The error is white TF32
                              noise plus isolated ReLU-gate flips where a pre-activation rounds
                              across zero (profiles/r2_grad_diag_1080p.log: no spatial structure)
    Adam style-loss trajectory relative, ceiling 3e-3
    Adam total-loss trajectory relative, ceiling 2e-2 -- except the 1080p case (9e-2): with content
                              init the content term of step 2 is ~1e-9, created entirely by the
                              first update, and is dominated by the gate-flip outliers
    Adam final image          relative L2, ceiling 2e-2
    timelapse frames          <= 3 LSB (truncating uint8 cast of values that differ by 1e-4)
L-BFGS without line search amplifies gradient differences: its third step is built from ONE
curvature pair whose y = g1 - g0 is a difference of nearly equal gradients.  ``lbfgs_noisy_64`` is a
well-conditioned start chosen so that the trajectory does not fork: it is pinned for all 8 steps;
the two older L-BFGS fixtures are pinned on their first two steps, and FusedLBFGS is pinned step by
step against torch.optim.LBFGS driven by the same gradients.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests import _cases as cases
from tests import _gpu_run

pytestmark = pytest.mark.gpu

ADAM_CASES = [n for n in cases.golden_names() if n.startswith("adam")]
LBFGS_CASES = [n for n in cases.golden_names() if n.startswith("lbfgs")]


# per-fixture gates = 1.5 x measured (profiles/r2_parity_report_v1.log), floors 3e-4 / 2e-3 (grad)
GATES: dict[str, dict[str, float]] = {
    "adam_clamp_64": {"layer_style_rel_max": 0.00051, "grad_rel_l2": 0.013, "total_rel_max": 0.0003, "style_rel_max": 0.0003, "final_rel_l2": 0.0001},
    "adam_content_1080p_c3": {"layer_style_rel_max": 0.0018, "grad_rel_l2": 0.023, "total_rel_max": 0.089, "style_rel_max": 0.00063, "final_rel_l2": 0.0001, "frames_max_lsb": 1, "frames_frac_diff": 0.001},
    "adam_content_256_c1": {"layer_style_rel_max": 0.0016, "grad_rel_l2": 0.023, "total_rel_max": 0.019, "style_rel_max": 0.0024, "final_rel_l2": 0.0025, "frames_max_lsb": 1, "frames_frac_diff": 0.13},
    "adam_content_512_c2": {"layer_style_rel_max": 0.0014, "grad_rel_l2": 0.02, "total_rel_max": 0.019, "style_rel_max": 0.00081, "final_rel_l2": 0.00054},
    "adam_content_64": {"layer_style_rel_max": 0.0011, "grad_rel_l2": 0.021, "total_rel_max": 0.0003, "style_rel_max": 0.00046, "final_rel_l2": 0.0014, "frames_max_lsb": 3, "frames_frac_diff": 0.024},
    "adam_layers_024_13_64": {"layer_style_rel_max": 0.0003, "grad_rel_l2": 0.011, "total_rel_max": 0.0003, "style_rel_max": 0.0003, "final_rel_l2": 0.0013},
    "adam_nonorm_64": {"layer_style_rel_max": 0.0017, "grad_rel_l2": 0.013, "total_rel_max": 0.004, "style_rel_max": 0.0018, "final_rel_l2": 0.00092, "frames_max_lsb": 2, "frames_frac_diff": 0.041},
    "adam_random_4k_c5": {"layer_style_rel_max": 0.00051, "grad_rel_l2": 0.034, "total_rel_max": 0.00032, "style_rel_max": 0.00037, "final_rel_l2": 0.00017},
    "adam_random_64": {"layer_style_rel_max": 0.00049, "grad_rel_l2": 0.022, "total_rel_max": 0.00043, "style_rel_max": 0.0003, "final_rel_l2": 0.02},
    "adam_stylesize_96x128": {"layer_style_rel_max": 0.0014, "grad_rel_l2": 0.02, "total_rel_max": 0.00042, "style_rel_max": 0.0003, "final_rel_l2": 0.00083},
    "adam_white_odd_70x94": {"layer_style_rel_max": 0.00047, "grad_rel_l2": 0.0058, "total_rel_max": 0.00033, "style_rel_max": 0.0003, "final_rel_l2": 0.0052},
    "lbfgs_content_64": {"layer_style_rel_max": 0.0011, "grad_rel_l2": 0.021},
    "lbfgs_random_64": {"layer_style_rel_max": 0.00049, "grad_rel_l2": 0.013},
    "lbfgs_noisy_64": {"layer_style_rel_max": 0.002, "grad_rel_l2": 0.02},
}
# lbfgs_noisy_64, per step: 1.5 x measured (profiles/r2_parity_lbfgs.log: 3e-4 3e-4 5.0e-2 7e-4 2e-4
# 8.2e-3 6.0e-3 4.3e-3); step 3 is the first step built from a curvature pair (see module docstring)
LBFGS_NOISY_RTOL = np.array([2e-3, 2e-3, 7.5e-2, 2e-3, 2e-3, 1.3e-2, 1.3e-2, 1.3e-2])
LBFGS_NOISY_FINAL = 1.1e-2   # measured 7.3e-3
CEILINGS = {"layer_style_rel_max": 2e-3, "grad_rel_l2": 3.5e-2, "total_rel_max": 2e-2,
            "style_rel_max": 3e-3, "final_rel_l2": 2e-2, "frames_max_lsb": 3,
            "frames_frac_diff": 0.15}
CEILING_EXCEPTIONS = {("adam_content_1080p_c3", "total_rel_max"): 9e-2}


def _gate(name: str, key: str) -> float:
    ceiling = CEILING_EXCEPTIONS.get((name, key), CEILINGS[key])
    return min(GATES.get(name, {}).get(key, ceiling), ceiling)


def _check_first_closure(name: str, m: dict[str, float]) -> None:
    assert m["layer_style_rel_max"] <= _gate(name, "layer_style_rel_max")
    assert m["layer_content_abs_over_total"] <= 1e-4
    assert m["grad_rel_l2"] <= _gate(name, "grad_rel_l2")
    assert m["grad_cosine"] >= 0.9995


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cudagraph"])
@pytest.mark.parametrize("name", ADAM_CASES)
def test_adam_case_matches_reference(name: str, graph: bool, cuda_device, tmp_path) -> None:  # noqa: ANN001, FBT001
    cfg, gold = cases.load_golden(name)
    csv_path = str(tmp_path / "loss.csv") if cfg.get("csv") else None
    res = _gpu_run.run_case(cfg, cuda_device, use_cuda_graph=graph, csv_path=csv_path)
    m = _gpu_run.compare(cfg, gold, res)
    _check_first_closure(name, m)
    assert len(res.total) == cfg["steps"]
    assert m["total_rel_max"] <= _gate(name, "total_rel_max")
    assert m["style_rel_max"] <= _gate(name, "style_rel_max")
    assert m["final_rel_l2"] <= _gate(name, "final_rel_l2")
    if cfg["init"] == "content":
        assert res.content[0] == 0.0  # targets come from the same kernels as the step forward
    if "frames_max_lsb" in m:
        assert len(res.frames) == cfg["steps"] // cfg["save_every"]
        assert res.frames[0].dtype == np.uint8
        assert res.frames[0].shape == (cfg["h"], cfg["w"], 3)
        assert m["frames_max_lsb"] <= _gate(name, "frames_max_lsb")
        assert m["frames_frac_diff"] <= _gate(name, "frames_frac_diff")
    if csv_path:
        lines = res.csv_text.strip().splitlines()
        assert lines[0] == "step,style_loss,content_loss,total_loss"
        assert len(lines) == cfg["steps"] + 1
        assert [ln.split(",")[0] for ln in lines[1:]] == [str(i) for i in range(1, cfg["steps"] + 1)]


@pytest.mark.parametrize("name", LBFGS_CASES)
def test_lbfgs_first_steps_match_reference(name: str, cuda_device) -> None:  # noqa: ANN001
    cfg, gold = cases.load_golden(name)
    res = _gpu_run.run_case(cfg, cuda_device, use_cuda_graph=False)
    m = _gpu_run.compare(cfg, gold, res)
    _check_first_closure(name, m)
    np.testing.assert_allclose(res.total[:2], gold["total_loss"][:2], rtol=2e-3)
    assert res.total[-1] < res.total[0]  # it optimises
    if name == "lbfgs_noisy_64":
        # the reference's DEFAULT optimiser pinned over the whole 8-step run, eagerly and through
        # the device-resident graph-captured step
        assert len(res.total) == 8
        err = np.abs(np.array(res.total) - gold["total_loss"]) / np.abs(gold["total_loss"])
        assert (err <= LBFGS_NOISY_RTOL).all(), err
        assert m["final_rel_l2"] <= LBFGS_NOISY_FINAL
        res_g = _gpu_run.run_case(cfg, cuda_device, use_cuda_graph=True)
        err_g = np.abs(np.array(res_g.total) - gold["total_loss"]) / np.abs(gold["total_loss"])
        assert (err_g <= LBFGS_NOISY_RTOL).all(), err_g


def test_fused_lbfgs_tracks_torch_lbfgs(cuda_device) -> None:  # noqa: ANN001
    """Same model, same gradients: FusedLBFGS must follow torch.optim.LBFGS step by step
    (default settings of the reference: lr=1, max_iter=1, max_eval=1, history 100)."""
    from style_transfer_visualizer_b200.optim import FusedLBFGS

    cfg, _gold = cases.load_golden("lbfgs_random_64")
    model, x0 = _gpu_run.build_model(cfg, cuda_device)
    runs = {}
    for kind in ("fused", "torch"):
        x = x0.clone().requires_grad_(True)
        opt = FusedLBFGS([x], lr=1.0, max_iter=1, max_eval=1) if kind == "fused" \
            else torch.optim.LBFGS([x], lr=1.0, max_iter=1, max_eval=1)
        losses: list[float] = []

        def closure(opt=opt, x=x, losses=losses):  # noqa: ANN001, ANN202
            opt.zero_grad()
            sl, cl = model(x)
            loss = cfg["style_w"] * torch.stack(sl).sum() + torch.stack(cl).sum()
            loss.backward()
            losses.append(float(loss.detach()))
            return loss

        for _ in range(10):
            opt.step(closure)
        runs[kind] = losses
    np.testing.assert_allclose(runs["fused"], runs["torch"], rtol=2e-2)
    np.testing.assert_allclose(runs["fused"][:4], runs["torch"][:4], rtol=1e-4)


def test_lbfgs_stationary_start_does_not_move(cuda_device) -> None:  # noqa: ANN001
    """torch's tolerance_grad early return (max|g| <= 1e-7): with content init and the default
    style weight the reference never moves the image on random weights."""
    from style_transfer_visualizer_b200.optim import FusedLBFGS

    cfg, _gold = cases.load_golden("adam_content_256_c1")
    model, x0 = _gpu_run.build_model(cfg, cuda_device)
    x = x0.clone().requires_grad_(True)
    opt = FusedLBFGS([x], lr=1.0, max_iter=1, max_eval=1)

    def closure():  # noqa: ANN202
        opt.zero_grad()
        sl, cl = model(x)
        loss = 1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()
        loss.backward()
        return loss

    opt.step(closure)
    assert float(x.grad.abs().max()) <= 1e-7
    assert torch.equal(x.detach(), x0)


def test_features_match_oracle_per_layer(cuda_device) -> None:  # noqa: ANN001
    """Tapped conv outputs against the CPU oracle, layer by layer (odd size: pools floor)."""
    from oracle import stv_oracle as orc

    cfg, _gold = cases.load_golden("adam_white_odd_70x94")
    model, _x0 = _gpu_run.build_model(cfg, cuda_device)
    content, _style, _init = cases.case_inputs(cfg)
    oracle = orc.OracleModel(orc.vgg19_features(cfg["weight_seed"]))
    want = oracle.taps(content)
    got = model.engine_for(cuda_device).tap_features_nchw(content.to(cuda_device))
    assert len(got) == len(want) == 6
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert cases.rel_l2(g.cpu().numpy(), w.numpy()) <= 3e-3


def test_model_surface_and_errors(cuda_device) -> None:  # noqa: ANN001
    """Structural pins of the reference's tests (tests/test_core_model.py:98-189)."""
    import style_transfer_visualizer_b200.core_model as cm

    cfg, _gold = cases.load_golden("adam_content_64")
    model, x0 = _gpu_run.build_model(cfg, cuda_device)
    sl, cl = model(x0.clone().requires_grad_(True))
    assert len(sl) == 5 and len(cl) == 1
    assert all(v.dim() == 0 for v in sl + cl)
    assert [t.shape for t in model.style_targets] == [(64, 64), (128, 128), (256, 256),
                                                      (512, 512), (512, 512)]
    assert model.content_targets[0].shape == (1, 512, 8, 8)
    g = cm.gram_matrix(torch.randn(1, 128, 9, 11, device=cuda_device))
    assert g.shape == (128, 128)
    assert torch.allclose(g, g.t(), atol=1e-6)
    assert torch.linalg.eigvalsh(g.double()).min() >= -1e-6
    # the reference's own Gram tests use a [1, 3, 100, 100] tensor (tests/test_core_model.py:33-96):
    # any channel count and batch folding must work, clamp included
    from oracle import stv_oracle as orc
    for shape, scale in [((1, 3, 100, 100), 1.0), ((2, 3, 17, 13), 1.0), ((1, 200, 6, 7), 1.0),
                         ((1, 3, 100, 100), 1e4)]:
        t = torch.randn(*shape, generator=torch.Generator().manual_seed(5)) * scale
        got = cm.gram_matrix(t.to(cuda_device)).cpu()
        want = orc.gram_matrix(t)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=2e-3, atol=2e-3 * float(want.abs().max()))
        assert float(got.max()) <= 5e5 / t[0].numel() / shape[0] * 1.0001
    model.style_targets = None
    with pytest.raises(RuntimeError, match="style_targets must be set"):
        model(x0)
    with pytest.raises(ValueError, match=r"\[1, 3, H, W\]"):
        model2, _ = _gpu_run.build_model(cfg, cuda_device)
        model2(torch.zeros(2, 3, 64, 64, device=cuda_device))


def test_large_size_properties(cuda_device) -> None:  # noqa: ANN001
    """BASELINE-size (512x512) checks that need no CPU oracle run: content init gives exactly zero
    content loss, the losses are reproducible bit-for-bit run to run (deterministic split-K), the
    gradient is linear in the loss weights, and a graph-replayed Adam run decreases the loss."""
    cfg = {"h": 512, "w": 512, "init": "content", "weight_seed": 0,
           "style_layers": [0, 5, 10, 19, 28], "content_layers": [21]}
    model, x0 = _gpu_run.build_model(cfg, cuda_device)

    def grad_for(sw: float, cw: float, x_in: torch.Tensor):  # noqa: ANN202
        x = x_in.clone().requires_grad_(True)
        sl, cl = model(x)
        (sw * torch.stack(sl).sum() + cw * torch.stack(cl).sum()).backward()
        return [float(v.detach()) for v in sl + cl], x.grad.clone()

    l1, g1 = grad_for(1e5, 1.0, x0)
    l2, g2 = grad_for(1e5, 1.0, x0)
    assert l1 == l2 and torch.equal(g1, g2)
    assert l1[-1] == 0.0
    noisy = x0 + 0.1 * torch.randn_like(x0)
    _, ga = grad_for(1e5, 0.0, noisy)
    _, gb = grad_for(0.0, 1.0, noisy)
    _, gab = grad_for(1e5, 1.0, noisy)
    assert cases.rel_l2((ga + gb).cpu().numpy(), gab.cpu().numpy()) <= 2e-3


@pytest.mark.parametrize("n,history", [(4099, 5), (30000, 100)])
def test_device_lbfgs_matches_torch_on_quadratic(n: int, history: int, cuda_device) -> None:  # noqa: ANN001
    """The device-resident step (coefficient-space recursion, ring eviction, odd length) against
    torch.optim.LBFGS on an ill-conditioned quadratic -- no TF32 anywhere, so the trajectories
    agree to fp32 rounding for many steps."""
    from style_transfer_visualizer_b200.optim import FusedLBFGS

    g = torch.Generator(device="cuda").manual_seed(7)
    diag = torch.logspace(-1, 1.5, n, device=cuda_device)
    b = torch.randn(n, device=cuda_device, generator=g)
    x0 = torch.randn(n, device=cuda_device, generator=g)
    runs = {}
    for kind in ("fused", "torch"):
        x = x0.clone().requires_grad_(True)
        opt = FusedLBFGS([x], lr=1.0, max_iter=1, max_eval=1, history_size=history) \
            if kind == "fused" else \
            torch.optim.LBFGS([x], lr=1.0, max_iter=1, max_eval=1, history_size=history)
        losses = []

        def closure(x=x, opt=opt, losses=losses):  # noqa: ANN001, ANN202
            opt.zero_grad()
            loss = 0.5 * (diag * x * x).sum() - (b * x).sum()
            loss.backward()
            losses.append(float(loss.detach()))
            return loss

        for _ in range(30):
            opt.step(closure)
        runs[kind] = (losses, x.detach().clone())
        if kind == "fused":
            counters = opt.device_counters()
            assert counters["n_iter"] == 30
            assert counters["pairs"] == min(history, 29)
    lf, lt = np.array(runs["fused"][0]), np.array(runs["torch"][0])
    scale = np.abs(lt).max()
    assert np.max(np.abs(lf - lt)) / scale < 2e-4
    assert lt[-1] < lt[0]
    assert cases.rel_l2(runs["fused"][1].cpu().numpy(), runs["torch"][1].cpu().numpy()) < 5e-3


def test_cli_end_to_end_offline(cuda_device, tmp_path, monkeypatch) -> None:  # noqa: ANN001
    """The reference's e2e CLI smoke (tests/test_cli.py:852-897: 64x64, 2 steps, --final-only,
    white init) through this package's CLI, with the offline random-weight switch, default
    optimiser (L-BFGS) and a TOML config + flag override."""
    from PIL import Image

    from style_transfer_visualizer_b200 import cli

    rng = np.random.default_rng(0)
    for name in ("content.png", "style.png"):
        Image.fromarray(rng.integers(0, 255, (64, 80, 3), dtype=np.uint8)).save(tmp_path / name)
    toml = tmp_path / "config.toml"
    toml.write_text("[optimization]\nsteps = 5\ninit_method = \"white\"\n[output]\nlog_every = 1\n")
    monkeypatch.setenv("STV_RANDOM_VGG_SEED", "0")
    out_dir = tmp_path / "out"
    rc = cli.main(["--content", str(tmp_path / "content.png"), "--style", str(tmp_path / "style.png"),
                   "--config", str(toml), "--steps", "2", "--final-only", "--output", str(out_dir),
                   "--log-loss", str(tmp_path / "loss.csv"), "--device", "cuda"])
    assert rc == 0
    png = out_dir / "stylized_content_x_style.png"
    assert png.exists()
    img = np.asarray(Image.open(png))
    assert img.shape == (64, 80, 3) and img.dtype == np.uint8
    rows = (tmp_path / "loss.csv").read_text().strip().splitlines()
    assert rows[0] == "step,style_loss,content_loss,total_loss" and len(rows) == 3


def test_1080p_step_properties(cuda_device) -> None:  # noqa: ANN001
    """BASELINE configs[2] size (1920x1080, odd pooled sizes 135 -> 67): content init gives exactly
    zero content loss, two evaluations are bit-identical, losses are finite and positive."""
    cfg = {"h": 1080, "w": 1920, "init": "content", "weight_seed": 0,
           "style_layers": [0, 5, 10, 19, 28], "content_layers": [21]}
    model, x0 = _gpu_run.build_model(cfg, cuda_device)
    outs = []
    for _ in range(2):
        x = x0.clone().requires_grad_(True)
        sl, cl = model(x)
        (1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()).backward()
        outs.append(([float(v.detach()) for v in sl + cl], x.grad.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])
    assert outs[0][0][-1] == 0.0
    assert all(v > 0 and np.isfinite(v) for v in outs[0][0][:-1])
    assert bool(torch.isfinite(outs[0][1]).all())


def test_job_runner_reuses_graph_and_matches_fresh_runs(cuda_device) -> None:  # noqa: ANN001
    """StyleJobRunner streams jobs through ONE captured graph (targets, image and Adam state are
    overwritten in place): every job must give the same result as a freshly built model + runner."""
    import style_transfer_visualizer_b200.core_model as cm
    from oracle import stv_oracle as orc
    from style_transfer_visualizer_b200 import jobs
    from style_transfer_visualizer_b200.config import StyleTransferConfig
    from style_transfer_visualizer_b200.optim import FusedAdam
    from style_transfer_visualizer_b200.optimization import OptimizationRunner

    def fresh_model():  # noqa: ANN202
        original = cm.initialize_vgg
        cm.initialize_vgg = lambda: orc.vgg19_features(0)
        try:
            return cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(cuda_device)
        finally:
            cm.initialize_vgg = original

    runner = jobs.StyleJobRunner(fresh_model(), 64, 96, steps=4, lr=0.01, device=cuda_device)
    pairs = [(orc.synthetic_image(10 + 2 * j, 64, 96), orc.synthetic_image(11 + 2 * j, 64, 96))
             for j in range(3)]
    results = [runner.run_job(c, s) for c, s in pairs]
    assert runner.jobs_done == 3
    graph_obj = runner._fused.graph  # noqa: SLF001
    for (content, style), (img, loss) in zip(pairs, results):
        model = fresh_model()
        model.set_targets(style.to(cuda_device), content.to(cuda_device))
        x = content.to(cuda_device).clone().requires_grad_(True)
        cfg = StyleTransferConfig.model_validate({
            "optimization": {"steps": 4, "style_w": 1e5, "content_w": 1.0, "lr": 0.01},
            "video": {"save_every": 5}, "output": {"log_every": 1}})
        _, hist, _ = OptimizationRunner(model, x, cfg, optimizer=FusedAdam([x], lr=0.01),
                                        progress_bar=_gpu_run.NullBar()).run()
        assert torch.equal(img.to(cuda_device), x.detach())
        assert loss == hist["total_loss"][-1]
    assert runner._fused.graph is graph_obj  # noqa: SLF001


def test_job_pool_lanes_match_sequential_runner(cuda_device) -> None:  # noqa: ANN001
    """StyleJobPool replays the step graphs of two independent jobs interleaved on two streams:
    every job's result must be bit-identical to the one-at-a-time StyleJobRunner's."""
    import style_transfer_visualizer_b200.core_model as cm
    from oracle import stv_oracle as orc
    from style_transfer_visualizer_b200 import jobs

    def fresh_model():  # noqa: ANN202
        original = cm.initialize_vgg
        cm.initialize_vgg = lambda: orc.vgg19_features(0)
        try:
            return cm.StyleContentModel([0, 5, 10, 19, 28], [21]).to(cuda_device)
        finally:
            cm.initialize_vgg = original

    pairs = [(orc.synthetic_image(30 + 2 * j, 64, 96), orc.synthetic_image(31 + 2 * j, 64, 96))
             for j in range(5)]
    seq = jobs.StyleJobRunner(fresh_model(), 64, 96, steps=5, lr=0.01, device=cuda_device)
    want = [seq.run_job(c, s) for c, s in pairs]
    pool = jobs.StyleJobPool(fresh_model, 64, 96, steps=5, lanes=2, lr=0.01, device=cuda_device)
    got = pool.run(pairs)
    assert len(got) == len(want)
    for (gi, gl), (wi, wl) in zip(got, want):
        assert torch.equal(gi, wi)
        assert gl == wl
    assert sum(r.jobs_done for r in pool.runners) == 5


def test_image_load_kernel_is_bit_identical_to_torchvision(cuda_device, tmp_path) -> None:  # noqa: ANN001
    """apply_transforms on the GPU ships bytes and normalises in one kernel: same bits as the
    reference's torchvision ToTensor + Normalize followed by .to(device) (image_io.py:64-84)."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms

    from style_transfer_visualizer_b200 import image_io
    from style_transfer_visualizer_b200.constants import IMAGENET_MEAN, IMAGENET_STD

    rng = np.random.default_rng(3)
    for h, w in [(64, 64), (70, 94), (129, 257)]:
        arr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        arr[0, :256 if w >= 256 else w, 0] = np.arange(min(w, 256), dtype=np.uint8)  # every byte value
        path = tmp_path / f"img_{h}x{w}.png"
        Image.fromarray(arr).save(path)
        for normalize in (False, True):
            got = image_io.load_image_to_tensor(str(path), cuda_device, normalize=normalize)
            pipeline = [transforms.ToTensor()]
            if normalize:
                pipeline.append(transforms.Normalize(mean=IMAGENET_MEAN, std=IMAGENET_STD))
            want = transforms.Compose(pipeline)(Image.open(path).convert("RGB")).unsqueeze(0)
            assert got.device.type == "cuda" and got.shape == want.shape
            assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cudagraph"])
def test_non_finite_warnings_are_deferred_to_the_log_cadence(graph: bool, cuda_device) -> None:  # noqa: ANN001, FBT001
    """Reference optimization.py:375-400: one warning per non-finite score per step.  On CUDA the
    flags are produced on the device (``stv_step_scores`` inside the captured graph, or
    ``stv_finite_flags`` for eager steps) and read back only when the host syncs for logging, in
    step order, with the reference's message texts -- no per-closure host sync."""
    import logging

    from style_transfer_visualizer_b200.config import StyleTransferConfig
    from style_transfer_visualizer_b200.logging_utils import logger
    from style_transfer_visualizer_b200.optim import FusedAdam
    from style_transfer_visualizer_b200.optimization import (OptimizationCallbacks,
                                                             OptimizationRunner)

    cfg, _gold = cases.load_golden("adam_content_64")
    model, x0 = _gpu_run.build_model(cfg, cuda_device)
    # poison one style target: style score and total become NaN at step 1; the NaN gradient then
    # makes the image (hence the content score) NaN from step 2 on
    model.engine_for(cuda_device).style_targets[1][0, 0] = float("nan")
    x = x0.clone().requires_grad_(True)
    config = StyleTransferConfig.model_validate({
        "optimization": {"steps": 6, "style_w": 1e5, "content_w": 1.0, "lr": 0.01},
        "video": {"save_every": 7}, "output": {"log_every": 3}})

    records: list[str] = []

    class Grab(logging.Handler):
        def emit(self, record: logging.LogRecord) -> None:
            if record.levelno >= logging.WARNING:
                records.append(record.getMessage())

    seen_at_step: list[int] = []
    handler = Grab()
    logger.addHandler(handler)
    try:
        runner = OptimizationRunner(
            model, x, config, optimizer=FusedAdam([x], lr=0.01), progress_bar=_gpu_run.NullBar(),
            callbacks=OptimizationCallbacks(on_step_end=lambda m: seen_at_step.append(len(records))),
            use_cuda_graph=graph)
        runner.run()
    finally:
        logger.removeHandler(handler)
    want = []
    for step in range(1, 7):
        want.append(f"Non-finite style score at step {step}")
        if step >= 2:
            want.append(f"Non-finite content score at step {step}")
        want.append(f"Non-finite total loss at step {step}, using previous loss")
    assert records == want
    # nothing is emitted between logging syncs: steps 1-2 see no warning yet, step 3 flushes 1..3
    assert seen_at_step[:2] == [0, 0]
    assert seen_at_step[2] == 2 + 3 + 3
    assert seen_at_step[3] == seen_at_step[4] == seen_at_step[2]
    assert seen_at_step[5] == len(want)


def test_models_release_their_gpu_memory(cuda_device) -> None:  # noqa: ANN001
    """An engine (packed weights, activation / gradient workspaces) lives exactly as long as the
    model that owns it: creating and dropping models must not grow the allocated GPU memory (the
    engine registry used by the custom ops holds weak references only)."""
    import gc

    import style_transfer_visualizer_b200.core_model as cm

    cfg, _gold = cases.load_golden("adam_content_64")

    def one_model() -> None:
        model, x0 = _gpu_run.build_model(cfg, cuda_device)
        x = x0.clone().requires_grad_(True)
        sl, cl = model(x)
        (1e5 * torch.stack(sl).sum() + torch.stack(cl).sum()).backward()
        assert model.engine_for(cuda_device).workspace_bytes(cfg["h"], cfg["w"]) > 0

    one_model()  # one-off allocations (cuBLAS-free path, but torch caches a few small buffers)
    gc.collect()
    torch.cuda.synchronize(cuda_device)
    base = torch.cuda.memory_allocated(cuda_device)
    registered = len(cm._ENGINES)  # noqa: SLF001
    for _ in range(3):
        one_model()
    gc.collect()
    torch.cuda.synchronize(cuda_device)
    assert len(cm._ENGINES) <= registered  # noqa: SLF001
    assert torch.cuda.memory_allocated(cuda_device) <= base + (1 << 20)
