"""Benchmark of the optimisation hot path (BASELINE.json metric: optimisation steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 512|1080p] [--impl reference]

One "step" = one full optimisation step of the reference's runner (VGG19 forward, Gram/MSE losses,
dgrad backward, Adam update of the image).  At N=1 the workload is BASELINE.json configs[1]
(VGG19 512x512, default style/content layers, Adam, random-init weights, synthetic images); with
N>1 (torchrun, one rank per GPU) every rank runs its own independent content/style pair
(configs[3]: independent jobs, no data-path collective => weak scaling).  Rank 0 prints ONE JSON
line.  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "512": dict(h=512, w=512, save_every=0, steps=300,
                name="configs[1]: VGG19 512x512, Adam, default style/content layers, "
                     "content init, random-init weights, synthetic images"),
    "1080p": dict(h=1080, w=1920, save_every=10, steps=100,
                  name="configs[2]: VGG19 1920x1080, Adam, save_every=10 frame readback, "
                       "content init, random-init weights, synthetic images"),
    "jobs64": dict(h=512, w=512, save_every=0, steps=30,
                   name="configs[3]: 64 independent 512x512 content/style pairs, Adam, "
                        "statically partitioned over the GPUs (no collective)"),
    "4k": dict(h=2160, w=3840, save_every=0, steps=20,
               name="configs[4]: VGG19 3840x2160 single image, Adam, row-band sharded conv stack "
                    "with halo exchange + Gram all-reduce over the GPUs"),
}
METRIC = "optimization steps/sec"
STYLE_W, CONTENT_W, LR = 1e5, 1.0, 0.01


def _peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(h: int, w: int, *, budget_s: float, max_steps: int, min_steps: int = 2):
    """Time the reference's algorithm (oracle port: same torch ops at the reference's call sites)
    on all host threads.  Returns (steps/sec, steps timed, threads)."""
    import torch

    from oracle import stv_oracle as orc

    torch.set_num_threads(os.cpu_count() or 1)
    content = orc.synthetic_image(1, h, w)
    style = orc.synthetic_image(2, h, w)
    model = orc.OracleModel(orc.vgg19_features(0))
    model.set_targets(style, content)
    x = orc.initialize_input(content, "content")
    opt = orc.make_optimizer("adam", x, LR)

    def closure():
        return orc.closure_step(model, x, STYLE_W, CONTENT_W)[2]

    opt.step(closure)  # warm-up (thread pools, oneDNN primitive caches)
    done, t0 = 0, time.perf_counter()
    while done < max_steps:
        opt.step(closure)
        done += 1
        if done >= min_steps and time.perf_counter() - t0 > budget_s:
            break
    elapsed = time.perf_counter() - t0
    return done / elapsed, done, torch.get_num_threads()


def run_reference_arm(args, wl) -> None:  # noqa: ANN001
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, done, threads = cpu_reference_run(wl["h"], wl["w"], budget_s=90.0,
                                             max_steps=max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": done, "warmup": 1, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["name"], "height": wl["h"], "width": wl["w"],
                   "optimizer": "adam", "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": f"{done} of {args.steps} requested steps at full resolution "
                                   "(90 s time budget), oracle port of the reference's path"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def _quiet_nccl() -> None:
    """stdout must carry exactly one JSON line: NCCL's banner (printed at NCCL_DEBUG >= VERSION) is
    sent to a file unless STV_NCCL_DEBUG asks for it."""
    if "STV_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["STV_NCCL_DEBUG"]
        return
    os.environ.pop("NCCL_DEBUG", None)
    os.environ["NCCL_DEBUG_FILE"] = os.path.join(tempfile.gettempdir(), "stv_nccl_%h_%p.log")
    os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)  # noqa: SIM115
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in Path(self.path).read_text().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        Path(self.path).unlink(missing_ok=True)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_job(device, h: int, w: int, seed_offset: int):  # noqa: ANN001
    """Model (random-init VGG19, seed 0) + pinned host images for one content/style pair."""
    import torch

    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import synthetic
    from style_transfer_visualizer_b200.constants import (DEFAULT_CONTENT_LAYERS,
                                                          DEFAULT_STYLE_LAYERS)

    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)  # same seed as the CPU arm
    try:
        model = cm.StyleContentModel(list(DEFAULT_STYLE_LAYERS),
                                     list(DEFAULT_CONTENT_LAYERS)).to(device)
    finally:
        cm.initialize_vgg = original
    content = synthetic.synthetic_image(1 + 2 * seed_offset, h, w).pin_memory()
    style = synthetic.synthetic_image(2 + 2 * seed_offset, h, w).pin_memory()
    torch.cuda.synchronize(device)
    return model, content, style


def profile_dominant_kernel(model, x, steps: int) -> dict:  # noqa: ANN001
    """CUDA-event timing of the tensor-core conv launches (3x3 fwd, 3x3 dgrad, 1x1 style-bwd) over
    `steps` eager executions of the step.  One event pair brackets every RUN of consecutive conv
    launches on the stream (a run ends at the next non-conv kernel: pool, loss, first layer ...), so
    launches that follow each other in the real step keep their programmatic-dependent-launch
    overlap, as they do in the graph-replayed timed region; an event pair around every single launch
    would add several microseconds of front-end latency to kernels that are only 15-50 us long at
    512x512.  Returns algorithmic FLOPs, summed run time and the launch count."""
    import torch

    from style_transfer_visualizer_b200 import _native as nat

    def conv_flops(name: str, a: tuple) -> float | None:
        if name in ("stv_conv3x3_fwd", "stv_conv3x3_fwd_bits", "stv_conv3x3_fwd_pool",
                    "stv_conv3x3_fwd_pool_code"):       # (x, w, bias, H, W, Cin, Cout, ...)
            return 2.0 * 9 * a[3] * a[4] * a[5] * a[6]
        if name in ("stv_conv3x3_dgrad", "stv_conv3x3_dgrad_bits", "stv_conv3x3_dgrad_unpool"):
            return 2.0 * 9 * a[2] * a[3] * a[4] * a[5]  # (dy, w, H, W, Cout, Cin, ...)
        if name == "stv_conv3x3_first_dgrad_tc":        # (dy, w16, H, W, Cout, ...): 3 real channels
            return 2.0 * 9 * a[2] * a[3] * a[4] * 3
        if name == "stv_style_bwd":                     # (x, s, hw, C, ...)
            return 2.0 * a[2] * a[3] * a[3]
        return None

    runs: list[tuple[torch.cuda.Event, torch.cuda.Event]] = []
    state = {"open": None, "flops": 0.0, "launches": 0}
    engine = model.engine_for(x.device)
    overlap = engine.overlap_losses
    engine.overlap_losses = False  # everything on one stream: a run is never overlapped by losses
    orig_call = nat.call

    def close_run() -> None:
        if state["open"] is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            runs.append((state["open"], e1))
            state["open"] = None

    def call(name: str, *a):  # noqa: ANN002, ANN202
        f = conv_flops(name, a)
        if f is None:
            close_run()
        else:
            if state["open"] is None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
                state["open"] = e0
            state["flops"] += f
            state["launches"] += 1
        return orig_call(name, *a)

    nat.call = call
    try:
        for _ in range(steps):
            # Eager Python launches are slower than the kernels at 512x512: a 40 ms spin kernel lets
            # the host queue the whole step first, so that the GPU never idles inside a run.
            torch.cuda._sleep(int(0.040 * 1.9e9))  # noqa: SLF001
            xx = x.detach().clone().requires_grad_(True)
            sl, cl = model(xx)
            (STYLE_W * torch.stack(sl).sum() + CONTENT_W * torch.stack(cl).sum()).backward()
            close_run()
        torch.cuda.synchronize()
    finally:
        nat.call = orig_call
        engine.overlap_losses = overlap
    ms = sum(e0.elapsed_time(e1) for e0, e1 in runs)
    return {"flops": state["flops"], "ms": ms, "launches": state["launches"], "steps": steps,
            "runs": len(runs)}


def run_gpu_arm(args, wl) -> None:  # noqa: ANN001, PLR0915
    import torch

    from style_transfer_visualizer_b200 import _native as nat
    from style_transfer_visualizer_b200 import jobs
    from style_transfer_visualizer_b200.config import StyleTransferConfig
    from style_transfer_visualizer_b200.optim import FusedAdam
    from style_transfer_visualizer_b200.optimization import OptimizationRunner

    _quiet_nccl()
    info = jobs.init_distributed()
    device = torch.device("cuda", info.local_rank)
    torch.cuda.set_device(device)
    nat.require_device(device)  # fails loudly if the .so is missing or the GPU is not sm_100
    h, w, k_steps, warm = wl["h"], wl["w"], args.steps, max(3, args.warmup)

    class Bar:
        def update(self, n=1): ...  # noqa: ANN001, ANN201
        def set_postfix(self, *a, **k): ...  # noqa: ANN002, ANN003, ANN201
        def close(self): ...  # noqa: ANN201

    class Sink:
        _size = None
        count = 0

        def append_data(self, f):  # noqa: ANN001, ANN201
            Sink.count += 1

        def close(self): ...  # noqa: ANN201

    def make_cfg(steps: int, log_every: int):  # noqa: ANN202
        return StyleTransferConfig.model_validate({
            "optimization": {"steps": steps, "style_w": STYLE_W, "content_w": CONTENT_W, "lr": LR,
                             "init_method": "content", "optimizer": "adam"},
            "video": {"save_every": wl["save_every"] or steps + 1},
            "output": {"log_every": log_every}})

    model, content_h, style_h = build_job(device, h, w, info.rank)

    # ---- device-resident measurement: inputs already in HBM when the timed region starts -------
    from style_transfer_visualizer_b200.core_model import initialize_input

    content = content_h.to(device, non_blocking=True)
    style = style_h.to(device, non_blocking=True)
    model.set_targets(style, content)
    x = initialize_input(content, "content")
    opt = FusedAdam([x], lr=LR)
    sink = Sink() if wl["save_every"] else None
    runner = OptimizationRunner(model, x, make_cfg(warm + k_steps, 10), optimizer=opt,
                                progress_bar=Bar(), video_writer=sink, use_cuda_graph=True)
    runner.prepare()

    def one_step() -> None:
        idx = runner._step_index + 1  # noqa: SLF001
        runner._active_step_idx = idx  # noqa: SLF001
        runner._fused_step(idx)  # noqa: SLF001
        runner._finalize_step(runner._pending_step_tensors)  # noqa: SLF001

    launches_before = nat.launch_count()
    for _ in range(warm):
        one_step()
    launches_per_step = runner._fused.kernel_launches  # noqa: SLF001
    torch.cuda.synchronize(device)
    jobs.barrier()
    sampler = ClockSampler(info.local_rank) if info.rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(k_steps):
        one_step()
    runner._drain_frames()  # noqa: SLF001
    e1.record()
    torch.cuda.synchronize(device)
    jobs.barrier()
    clocks = sampler.stop() if sampler else None
    ms_local = e0.elapsed_time(e1)
    ms_total = jobs.max_over_ranks(ms_local, device)
    value = info.world_size * k_steps / (ms_total / 1e3)
    del launches_before

    # ---- end to end through the public API: pinned host inputs, H2D + per-step loss D2H -------
    e2e_steps = k_steps

    def e2e_pass():  # noqa: ANN202
        torch.cuda.synchronize(device)
        jobs.barrier()
        t0 = time.perf_counter()
        c_dev = content_h.to(device, non_blocking=True)
        s_dev = style_h.to(device, non_blocking=True)
        model.set_targets(s_dev, c_dev)
        x2 = initialize_input(c_dev, "content")
        runner2 = OptimizationRunner(model, x2, make_cfg(e2e_steps, 1),
                                     optimizer=FusedAdam([x2], lr=LR), progress_bar=Bar(),
                                     video_writer=Sink() if wl["save_every"] else None)
        t1 = time.perf_counter()
        runner2.prepare()
        torch.cuda.synchronize(device)
        t2 = time.perf_counter()
        final, hist, _ = runner2.run()
        final_host = final.detach().cpu()
        torch.cuda.synchronize(device)
        t3 = time.perf_counter()
        phases = {"h2d_targets_s": t1 - t0, "prepare_graph_capture_s": t2 - t1, "run_s": t3 - t2}
        total = jobs.max_over_ranks(time.perf_counter() - t0, device)
        return total, phases, hist, final_host, runner2

    # Two complete passes, the faster one is reported (host-side jitter on a shared box moves this
    # wall-clock number by 2x between otherwise identical runs; both totals are kept in `phases`).
    passes = [e2e_pass() for _ in range(2)]
    e2e_s, e2e_phases, hist, final_host, runner2 = min(passes, key=lambda r: r[0])
    e2e_phases["pass_totals_s"] = [r[0] for r in passes]
    assert len(hist["total_loss"]) == e2e_steps
    h2d = (content_h.numel() + style_h.numel()) * 4 / e2e_steps
    d2h = 12 + final_host.numel() * 4 / e2e_steps + runner2.frame_bytes_d2h / e2e_steps
    e2e_value = info.world_size * e2e_steps / e2e_s

    if info.rank != 0:
        jobs.barrier()  # rank 0 finishes its single-rank extras (roofline pass) first
        jobs.shutdown()
        return

    # ---- roofline of the dominant kernel (tensor-core conv), measured live with CUDA events ----
    prof = profile_dominant_kernel(model, x, 3)
    peaks = _peaks()
    tf32_peak = peaks["bf16_sustained"] / 2.0
    achieved = prof["flops"] / (prof["ms"] / 1e3) / 1e12
    traffic = None
    rf = ROOT / "profiles" / "roofline_traffic.json"
    if rf.exists():
        traffic = json.loads(rf.read_text()).get(args.workload)
    step_flops = model.engine_for(device).flops_per_step(h, w)
    roofline = {
        "bound": "tensor", "kernel": "conv_igemm_tf32_kernel (3x3 fwd + dgrad, 1x1 style-bwd)",
        "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
        "traffic": traffic,
        "peak_source": f"{peaks['source']} bf16_tflops_sustained / 2 (TF32 tensor rate is half "
                       "the bf16 rate)",
        "launches_per_step": prof["launches"] // prof["steps"],
        "avg_launch_ms": prof["ms"] / prof["launches"],
        "algorithmic_flops_per_step": prof["flops"] / prof["steps"],
        "share_of_step_time": (prof["ms"] / prof["steps"]) / (ms_total / k_steps),
        "whole_step_tflops": step_flops["total_tri"] / (ms_total / k_steps / 1e3) / 1e12,
        "event_pairs_per_step": prof["runs"] // prof["steps"],
        "measured_on": "3 eager executions of the same step, each queued behind a 40 ms spin kernel "
                       "so that host launch latency is not counted as kernel time; one CUDA-event "
                       "pair on the launching stream around every run of consecutive conv launches "
                       "(runs end at the next pool / loss / first-layer kernel)",
    }

    # ---- CPU baseline (bounded sample, rank 0, N=1 only) ---------------------------------------
    cpu = None
    if info.world_size == 1 and not args.no_cpu_baseline:
        v, done, threads = cpu_reference_run(h, w, budget_s=12.0, max_steps=20)
        cpu = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port",
               "sample": f"{done} Adam steps at {w}x{h} (12 s budget) of the oracle port of the "
                         "reference's path, all host threads"}

    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": info.world_size,
        "steps": k_steps, "warmup": warm, "ms_per_step": ms_total / k_steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
        "data": "synthetic",
        "config": {"workload": wl["name"], "height": h, "width": w, "optimizer": "adam",
                   "lr": LR, "style_w": STYLE_W, "content_w": CONTENT_W,
                   "jobs_per_gpu": 1, "parallelism": f"independent jobs x{info.world_size}",
                   "l2": "no flush: each step streams ~%.1f GB of activations/gradients, far above "
                         "the 126 MB L2" % (model.engine_for(device).workspace_bytes(h, w) / 1e9),
                   "storage": "fp32 activations, TF32 multiply / FP32 accumulate"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "phases": e2e_phases,
                "how": "pinned host images -> H2D -> set_targets -> OptimizationRunner.run() with "
                       "log_every=1 (loss D2H every step) -> final image D2H; faster of two "
                       "complete passes"},
        "gpu_launches": launches_per_step * k_steps,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "frames_read_back": Sink.count,
    }
    print(json.dumps(line), flush=True)
    jobs.barrier()
    jobs.shutdown()


def run_jobs_arm(args, wl) -> None:  # noqa: ANN001
    """configs[3]: 64 independent jobs; every rank runs its static share through StyleJobRunner
    (model, workspaces, pinned buffers and the captured step graph reused across jobs).  The timed
    region is end to end per job: pinned host images in, result image out."""
    import torch

    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import _native as nat
    from style_transfer_visualizer_b200 import jobs, synthetic

    _quiet_nccl()
    info = jobs.init_distributed()
    device = torch.device("cuda", info.local_rank)
    torch.cuda.set_device(device)
    nat.require_device(device)
    n_jobs, steps = 64, args.steps
    mine = jobs.partition_jobs(n_jobs, info.world_size, info.rank)
    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
    try:
        model = cm.StyleContentModel([0, 5, 10, 19, 28], [21])
    finally:
        cm.initialize_vgg = original
    lanes = max(1, args.lanes)

    def make_model():  # noqa: ANN202
        cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
        try:
            return cm.StyleContentModel([0, 5, 10, 19, 28], [21])
        finally:
            cm.initialize_vgg = original

    models = iter([model] + [make_model() for _ in range(lanes - 1)])
    pool = jobs.StyleJobPool(lambda: next(models), wl["h"], wl["w"], steps=steps, lanes=lanes,
                             lr=LR, style_w=STYLE_W, content_w=CONTENT_W, device=device)
    pairs = [(synthetic.synthetic_image(1 + 2 * j, wl["h"], wl["w"]),
              synthetic.synthetic_image(2 + 2 * j, wl["h"], wl["w"])) for j in mine]
    pool.run(pairs[:lanes])  # warm-up wave: workspace allocation + graph capture in every lane
    torch.cuda.synchronize(device)
    jobs.barrier()
    t0 = time.perf_counter()
    losses = [r[1] for r in pool.run(pairs)]
    torch.cuda.synchronize(device)
    local = time.perf_counter() - t0
    jobs.barrier()
    total_s = jobs.max_over_ranks(local, device)
    if info.rank == 0:
        line = {
            "metric": METRIC, "value": n_jobs * steps / total_s, "unit": "steps/s",
            "n_gpus": info.world_size, "steps": steps, "warmup": 1,
            "ms_per_step": total_s / (len(mine) * steps) * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": wl["name"], "jobs": n_jobs, "jobs_per_gpu": len(mine),
                       "concurrent_lanes_per_gpu": lanes,
                       "steps_per_job": steps, "height": wl["h"], "width": wl["w"],
                       "timing": "host wall clock per rank over its jobs (H2D of both images and "
                                 "D2H of the result inside), max over ranks"},
            "jobs_per_second": n_jobs / total_s,
            "e2e": {"value": n_jobs * steps / total_s, "unit": "steps/s",
                    "h2d_bytes_per_step": 2 * 3 * wl["h"] * wl["w"] * 4 / steps,
                    "d2h_bytes_per_step": 3 * wl["h"] * wl["w"] * 4 / steps},
            "final_loss_first_job": losses[0],
        }
        print(json.dumps(line), flush=True)
    jobs.barrier()
    jobs.shutdown()


def run_sharded_arm(args, wl) -> None:  # noqa: ANN001
    """configs[4]: ONE image split into row bands over all ranks (strong scaling)."""
    import torch

    from style_transfer_visualizer_b200 import _native as nat
    from style_transfer_visualizer_b200 import jobs, synthetic
    from style_transfer_visualizer_b200.optim import FusedAdam
    from style_transfer_visualizer_b200.sharded import ShardedStyleContentModel

    _quiet_nccl()
    info = jobs.init_distributed()
    if info.world_size == 1 and not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        torch.distributed.init_process_group("nccl", rank=0, world_size=1,
                                             device_id=torch.device("cuda", 0))
    device = torch.device("cuda", info.local_rank)
    torch.cuda.set_device(device)
    nat.require_device(device)
    h, w, k_steps, warm = wl["h"], wl["w"], args.steps, max(3, args.warmup)
    model = ShardedStyleContentModel(synthetic.random_vgg19_features(0), [0, 5, 10, 19, 28], [21],
                                     device)
    content = synthetic.synthetic_image(1, h, w)
    style = synthetic.synthetic_image(2, h, w)
    model.set_targets(style, content)
    from style_transfer_visualizer_b200.sharded import ShardedFusedStep

    x = model.band_of(content)
    fused = ShardedFusedStep(model, x, lr=LR, style_w=STYLE_W, content_w=CONTENT_W)
    last = []
    del FusedAdam

    class _Opt:
        @staticmethod
        def step(_closure):  # noqa: ANN001, ANN205
            last[:] = [fused.step()]

    opt, closure = _Opt, None
    launches0 = nat.launch_count()
    for _ in range(warm):
        opt.step(closure)
    per_step = (nat.launch_count() - launches0) // warm
    graph_mode = "eager"
    if args.sharded_graph:
        fused.capture()
        graph_mode = "cuda-graph (NCCL send/recv + all-reduce captured)"
        for _ in range(2):
            opt.step(closure)
    torch.cuda.synchronize(device)
    jobs.barrier()
    sampler = ClockSampler(info.local_rank) if info.rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k_steps):
        opt.step(closure)
    e1.record()
    torch.cuda.synchronize(device)
    jobs.barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = jobs.max_over_ranks(e0.elapsed_time(e1), device)
    loss = float(last[0])
    if info.rank == 0:
        eng = model.engine.base
        flops = eng.flops_per_step(h, w)["total_tri"]
        line = {
            "metric": METRIC, "value": k_steps / (ms_total / 1e3), "unit": "steps/s",
            "n_gpus": info.world_size, "steps": k_steps, "warmup": warm,
            "ms_per_step": ms_total / k_steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": wl["name"], "height": h, "width": w, "optimizer": "adam",
                       "parallelism": f"row bands x{info.world_size} (16-row aligned), 1 halo row "
                                      "send/recv per conv, one Gram all-reduce per step",
                       "l2": "no flush: working set far above the 126 MB L2",
                       "step_launch": graph_mode},
            "e2e": None, "gpu_launches": per_step * k_steps, "gpu_launches_per_step": per_step,
            "clocks": clocks, "final_loss": loss,
            "whole_job_tflops": flops / (ms_total / k_steps / 1e3) / 1e12,
        }
        print(json.dumps(line), flush=True)
    # Tear-down: drop the captured graph before the communicator goes away, and leave without
    # destroy_process_group() -- destroying an NCCL communicator that was used inside a captured
    # graph blocked at exit on this stack (torch 2.11 / NCCL 2.28).
    torch.cuda.synchronize(device)
    fused.graph = None
    import gc

    gc.collect()
    jobs.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="512", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compact-backward", type=int, default=1,
                    help="0 = fp32 re-reads for the ReLU / pool backward (A/B against the bit codes)")
    ap.add_argument("--lanes", type=int, default=2,
                    help="jobs64 workload: independent jobs in flight per GPU (own stream + graph)")
    ap.add_argument("--sharded-graph", type=int, default=1,
                    help="4k workload: replay the sharded step from a CUDA graph (0 = eager)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if not args.compact_backward:
        import style_transfer_visualizer_b200.engine as _eng

        _eng.DEFAULT_COMPACT_BACKWARD = False
    if args.steps is None:
        args.steps = wl["steps"]
    if args.impl == "reference":
        run_reference_arm(args, wl)
    elif args.workload == "4k":
        run_sharded_arm(args, wl)
    elif args.workload == "jobs64":
        run_jobs_arm(args, wl)
    else:
        run_gpu_arm(args, wl)


if __name__ == "__main__":
    main()
