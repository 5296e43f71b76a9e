"""Benchmark of the optimisation hot path (BASELINE.json metric: optimisation steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 512|1080p|jobs64|4k]
                    [--impl reference] [--no-extras] [--no-cpu-baseline]

One "step" = one full optimisation step of the reference's runner (VGG19 forward, Gram/MSE losses,
dgrad backward, Adam update of the image).  At N=1 the workload is BASELINE.json configs[1]
(VGG19 512x512, default style/content layers, Adam, random-init weights, synthetic images); with
N>1 (torchrun, one rank per GPU) every rank runs its own independent content/style pair
(configs[3]: independent jobs, no data-path collective => weak scaling).  Rank 0 prints ONE JSON
line.  The same line carries, under "workloads", the other BASELINE configs measured in the same
process: 1920x1080 with save_every=10 (configs[2]) and the 64-job batch (configs[3]) at N=1, the
row-band sharded 3840x2160 image (configs[4]) at N>=2, and the reference's algorithm through stock
PyTorch (cuDNN / cuBLAS) on the same GPU.  See DESIGN.md "Measurement" for how each field is
obtained.

Timing: after W warm-up steps and a >= 0.5 s warm run, blocks of exactly K steps are timed with
CUDA events (barrier + synchronize on both sides of every block, max over ranks) until at least
5 blocks and 1 s have been measured; the MEDIAN block is reported (all block times are kept in
"blocks_ms").  Clocks and throttle reasons are sampled through NVML every 20 ms during those
blocks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
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
STYLE_LAYERS, CONTENT_LAYERS = [0, 5, 10, 19, 28], [21]


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
def cpu_reference_run(h: int, w: int, *, budget_s: float, max_steps: int, min_steps: int = 2,
                      warmup: int = 1):
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

    for _ in range(max(1, warmup)):  # thread pools, oneDNN primitive caches
        opt.step(closure)
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
    # a bounded sample: at most `--steps` steps and ~90 s; warm-up capped so that the slow 1080p /
    # 4K configurations stay within minutes
    warm = max(1, min(args.warmup, 5 if wl["h"] * wl["w"] <= 512 * 512 else 1))
    value, done, threads = cpu_reference_run(wl["h"], wl["w"], budget_s=90.0,
                                             max_steps=max(1, args.steps), warmup=warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "warmup_run": warm,
        "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["name"], "height": wl["h"], "width": wl["w"],
                   "optimizer": "adam", "lr": LR, "style_w": STYLE_W, "content_w": CONTENT_W,
                   "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": f"{done} of {args.steps} requested steps at full resolution "
                                   f"after {warm} warm-up steps (90 s time budget), oracle port of "
                                   "the reference's path on all host threads"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled through NVML from a background thread every
    20 ms (an `nvidia-smi -lms` subprocess needs longer to start than a short timed region lasts).
    ``window(t0, t1)`` summarises the samples whose host timestamp falls inside a timed region."""

    REASONS = ((0x4, "sw_power_cap"), (0x8, "hw_slowdown"), (0x20, "sw_thermal_slowdown"),
               (0x40, "hw_thermal_slowdown"))

    def __init__(self, cuda_index: int, period_s: float = 0.02) -> None:
        self.period = period_s
        self.samples: list[tuple[float, float, int, float]] = []
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None
        self.max_mhz: float | None = None
        self.error: str | None = None
        self._handle = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                self._handle = pynvml.nvmlDeviceGetHandleByUUID(
                    (uuid if uuid.startswith("GPU-") else "GPU-" + uuid).encode())
            except Exception:  # noqa: BLE001
                self._handle = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle,
                                                                  pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001
            self.error = f"NVML unavailable: {type(exc).__name__}: {exc}"

    def _loop(self) -> None:
        nv, h = self._nv, self._handle
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                rs = int(reasons_fn(h))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception as exc:  # noqa: BLE001
                self.error = f"{type(exc).__name__}: {exc}"
                return
            self._stop.wait(self.period)

    def start(self) -> None:
        if self._handle is None:
            return
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def window(self, t0: float, t1: float) -> dict:
        rows = [s for s in self.samples if t0 <= s[0] <= t1]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": self.error or "no NVML sample inside the timed region"}
        mask = 0
        for r in rows:
            mask |= r[2]
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz,
                "reasons": [name for bit, name in self.REASONS if mask & bit],
                "samples": len(rows), "power_w_max": max(r[3] for r in rows),
                "sampled_with": "NVML, 20 ms period, inside the timed blocks"}


# ------------------------------------------------------------------------------------------------
# GPU arm helpers
# ------------------------------------------------------------------------------------------------
class Bar:
    def update(self, n=1): ...  # noqa: ANN001, ANN201
    def set_postfix(self, *a, **k): ...  # noqa: ANN002, ANN003, ANN201
    def close(self): ...  # noqa: ANN201


class Sink:
    _size = None

    def __init__(self) -> None:
        self.count = 0

    def append_data(self, f):  # noqa: ANN001, ANN201
        self.count += 1

    def close(self): ...  # noqa: ANN201


def make_cfg(steps: int, log_every: int, save_every: int):  # noqa: ANN201
    from style_transfer_visualizer_b200.config import StyleTransferConfig

    return StyleTransferConfig.model_validate({
        "optimization": {"steps": steps, "style_w": STYLE_W, "content_w": CONTENT_W, "lr": LR,
                         "init_method": "content", "optimizer": "adam"},
        "video": {"save_every": save_every or steps + 1},
        "output": {"log_every": log_every}})


def build_job(device, h: int, w: int, seed_offset: int):  # noqa: ANN001, ANN201
    """Model (random-init VGG19, seed 0) + pinned host images for one content/style pair."""
    import torch

    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import synthetic

    original = cm.initialize_vgg
    cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)  # same seed as the CPU arm
    try:
        model = cm.StyleContentModel(list(STYLE_LAYERS), list(CONTENT_LAYERS)).to(device)
    finally:
        cm.initialize_vgg = original
    content = synthetic.synthetic_image(1 + 2 * seed_offset, h, w).pin_memory()
    style = synthetic.synthetic_image(2 + 2 * seed_offset, h, w).pin_memory()
    torch.cuda.synchronize(device)
    return model, content, style


CONV_FWD = ("stv_conv3x3_fwd", "stv_conv3x3_fwd_bits", "stv_conv3x3_fwd_pool",
            "stv_conv3x3_fwd_pool_code")
CONV_DGRAD = ("stv_conv3x3_dgrad", "stv_conv3x3_dgrad_bits", "stv_conv3x3_dgrad_unpool")


def conv_flops(name: str, a: tuple) -> float | None:
    """Algorithmic FLOPs of one tensor-core conv launch (every tcgen05 conv entry point)."""
    if name in CONV_FWD:                            # (x, w, bias, H, W, Cin, Cout, ...)
        return 2.0 * 9 * a[3] * a[4] * a[5] * a[6]
    if name in CONV_DGRAD:                          # (dy, w, H, W, Cout, Cin, ...)
        return 2.0 * 9 * a[2] * a[3] * a[4] * a[5]
    if name == "stv_conv3x3_dgrad_bits_style":      # dgrad + the fused 1x1 Gram backward of the layer
        return 2.0 * 9 * a[2] * a[3] * a[4] * a[5] + 2.0 * a[2] * a[3] * a[5] * a[5]
    if name in ("stv_conv3x3_first_dgrad_tc", "stv_conv3x3_first_dgrad_rows"):  # (dy, w, H, W, Cout, ...): 3 real channels
        return 2.0 * 9 * a[2] * a[3] * a[4] * 3
    if name == "stv_style_bwd":                     # (x, s, hw, C, ...)
        return 2.0 * a[2] * a[3] * a[3]
    return None


def profile_dominant_kernel(model, x, steps: int) -> dict:  # noqa: ANN001
    """CUDA-event timing of EVERY tensor-core conv launch of the step (3x3 fwd incl. the pool-fused
    ones, 3x3 dgrad incl. the un-pooling ones, the first-layer dgrad, the 1x1 style backward)
    over `steps` eager executions.  One event pair brackets every RUN of consecutive conv launches
    on the stream (a run ends at the next non-conv kernel), so launches that follow each other in
    the real step keep their programmatic-dependent-launch overlap, as they do in the graph-replayed
    timed region; an event pair around every single launch would add several microseconds of
    front-end latency to kernels that are only 15-50 us long at 512x512."""
    import torch

    from style_transfer_visualizer_b200 import _native as nat

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


def time_blocks(step_fn, k_steps: int, device, *, end_of_block=None, min_blocks: int = 5,  # noqa: ANN001
                min_total_s: float = 1.0, max_blocks: int = 60, warm_s: float = 0.5):
    """Warm run of >= warm_s, then blocks of exactly k_steps timed with CUDA events (barrier +
    synchronize on both sides, max over ranks).  Returns (block times in ms, host t0, host t1)."""
    import torch

    from style_transfer_visualizer_b200 import jobs

    t_w = time.perf_counter()
    while True:  # clocks up, caches and allocator settled
        for _ in range(k_steps):
            step_fn()
        torch.cuda.synchronize(device)
        done = jobs.max_over_ranks(1.0 if time.perf_counter() - t_w >= warm_s else 0.0, device)
        if done >= 1.0:
            break
    if end_of_block:
        end_of_block()
    blocks: list[float] = []
    host_t0 = time.perf_counter()
    while len(blocks) < min_blocks or (sum(blocks) < min_total_s * 1e3 and len(blocks) < max_blocks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        jobs.barrier()
        e0.record()
        for _ in range(k_steps):
            step_fn()
        if end_of_block:
            end_of_block()
        e1.record()
        torch.cuda.synchronize(device)
        jobs.barrier()
        blocks.append(jobs.max_over_ranks(e0.elapsed_time(e1), device))
    return blocks, host_t0, time.perf_counter()


def tensor_peak(peaks: dict, clocks: dict | None) -> tuple[float, str]:
    """TF32 tensor-pipe denominator by regime: the burst figure when the timed region ran at full
    clock without a power cap, the sustained one when sw_power_cap was sampled (TF32 rate = half
    the measured bf16 rate)."""
    capped = bool(clocks and "sw_power_cap" in (clocks.get("reasons") or []))
    slow = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz")
                and clocks["sm_mhz"] < 0.9 * clocks["sm_max_mhz"])
    if capped or slow:
        return peaks["bf16_sustained"] / 2.0, (
            f"{peaks['source']} bf16_tflops_sustained / 2 (sw_power_cap or reduced SM clock "
            "sampled during the timed blocks; TF32 tensor rate is half the bf16 rate)")
    return peaks["bf16_burst"] / 2.0, (
        f"{peaks['source']} bf16_tflops (burst) / 2 (full SM clock, no power cap during the "
        "timed blocks; TF32 tensor rate is half the bf16 rate)")


def hbm_kernel_rooflines(model, x, device, peaks: dict) -> list[dict]:  # noqa: ANN001
    """Achieved GB/s of the memory-bound kernels on this workload's buffers: algorithmic bytes /
    CUDA-event time, each kernel alone and hot, 10 launches back to back (buffers > L2 only at
    1080p and above; at 512x512 several of them fit the 126 MB L2 and the entry says so)."""
    import torch

    from style_transfer_visualizer_b200 import ops

    eng = model.engine_for(device)
    h, w = int(x.shape[2]), int(x.shape[3])
    ws = eng._workspace(h, w, with_grad=True)  # noqa: SLF001
    out: list[dict] = []

    def timed(name: str, nbytes: float, fn) -> None:  # noqa: ANN001
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / 10
        gbs = nbytes / ms / 1e6
        out.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes": nbytes,
                    "ms": ms, "fits_l2": nbytes < 126e6})

    n = x.numel()
    g = ws.grad_img
    m, v = torch.zeros_like(x), torch.zeros_like(x)
    xa = x.detach().clone()
    timed("adam_step_kernel", 28.0 * n,
          lambda: ops.adam_step(xa, g, m, v, beta1=0.9, beta2=0.999, eps=1e-8, step_size=1e-3,
                                bias2_sqrt=1.0))
    st0 = eng.stages[0]
    timed("conv_first_fwd_kernel (+sign bits)", 4.0 * n + 2 * 4.0 * 64 * h * w + 8.0 * h * w,
          lambda: ops.conv3x3_first_fwd(x.detach(), st0.weight, st0.bias, ws.pre[0], ws.post[0],
                                        out_bits=ws.bits[0]))
    timed("conv_first_dgrad_tc_kernel (first-layer dgrad)", 4.0 * 64 * h * w + 4.0 * n,
          lambda: ops.conv3x3_first_dgrad_rows(ws.d_y[0], st0.w_dgrad, g))
    k0 = 0
    t0 = eng._tap_tensor(ws, eng.style_idx[k0])  # noqa: SLF001
    timed("gram_partial + finalize (C=64)", 4.0 * t0.numel(),
          lambda: ops.gram_loss_fwd(t0, ws.gram_ws[k0], target=eng.style_targets[k0],
                                    s_out=ws.s_mat[k0], loss_out=ws.losses[k0:k0 + 1]))
    ci = eng.content_idx[0]
    tc = eng._tap_tensor(ws, ci)  # noqa: SLF001
    tgt = eng.content_targets_nhwc[0]
    ns = len(eng.style_idx)
    timed("sqdiff_partial (content MSE fwd)", 8.0 * tc.numel(),
          lambda: ops.content_loss_fwd(tc, tgt, ws.scratch, ws.losses[ns:ns + 1]))
    u8 = torch.empty(h, w, 3, device=device, dtype=torch.uint8)
    timed("frame_to_u8_kernel", 15.0 * h * w,
          lambda: ops.frame_to_u8(x.detach(), u8, denormalize=True))
    return out


def measure_single_image(wl_key: str, args, info, device, sampler, *, full: bool) -> dict:  # noqa: ANN001
    """steps/s of one image size: device-resident graph-replayed steps (`value`), the public API
    end to end from pinned host buffers (`e2e`), and the roofline of the conv kernels."""
    import torch

    from style_transfer_visualizer_b200 import jobs
    from style_transfer_visualizer_b200.core_model import initialize_input
    from style_transfer_visualizer_b200.optim import FusedAdam
    from style_transfer_visualizer_b200.optimization import OptimizationRunner

    wl = WORKLOADS[wl_key]
    h, w, k_steps, warm = wl["h"], wl["w"], args.steps, max(3, args.warmup)
    model, content_h, style_h = build_job(device, h, w, info.rank)

    # ---- device-resident measurement: inputs already in HBM when the timed region starts -------
    content = content_h.to(device, non_blocking=True)
    style = style_h.to(device, non_blocking=True)
    model.set_targets(style, content)
    x = initialize_input(content, "content")
    sink = Sink() if wl["save_every"] else None
    runner = OptimizationRunner(model, x, make_cfg(10 ** 7, 10, wl["save_every"]),
                                optimizer=FusedAdam([x], lr=LR), progress_bar=Bar(),
                                video_writer=sink, use_cuda_graph=True)
    runner.prepare()

    def one_step() -> None:
        idx = runner._step_index + 1  # noqa: SLF001
        runner._active_step_idx = idx  # noqa: SLF001
        runner._fused_step(idx)  # noqa: SLF001
        runner._finalize_step(runner._pending_step_tensors)  # noqa: SLF001

    for _ in range(warm):
        one_step()
    launches_per_step = runner._fused.kernel_launches  # noqa: SLF001
    blocks, t0, t1 = time_blocks(one_step, k_steps, device, end_of_block=runner._drain_frames,  # noqa: SLF001
                                 min_blocks=5 if full else 3,
                                 min_total_s=1.0 if full else 0.6)
    ms_block = statistics.median(blocks)
    clocks = sampler.window(t0, t1) if sampler else None
    value = info.world_size * k_steps / (ms_block / 1e3)
    frames = sink.count if sink else 0
    steps_done = runner._step_index  # noqa: SLF001

    # ---- end to end through the public API: pinned host inputs, H2D + per-step loss D2H -------
    final_pinned = torch.empty(1, 3, h, w, dtype=torch.float32, pin_memory=True)

    def e2e_pass():  # noqa: ANN202
        torch.cuda.synchronize(device)
        jobs.barrier()
        t_0 = time.perf_counter()
        c_dev = content_h.to(device, non_blocking=True)
        s_dev = style_h.to(device, non_blocking=True)
        model.set_targets(s_dev, c_dev)
        x2 = initialize_input(c_dev, "content")
        snk = Sink() if wl["save_every"] else None
        runner2 = OptimizationRunner(model, x2, make_cfg(k_steps, 1, wl["save_every"]),
                                     optimizer=FusedAdam([x2], lr=LR), progress_bar=Bar(),
                                     video_writer=snk)
        t_1 = time.perf_counter()
        runner2.prepare()
        torch.cuda.synchronize(device)
        t_2 = time.perf_counter()
        final, hist, _ = runner2.run()
        final_host = final_pinned.copy_(final.detach(), non_blocking=True)  # result image -> pinned host
        torch.cuda.synchronize(device)
        t_3 = time.perf_counter()
        phases = {"h2d_targets_s": t_1 - t_0, "prepare_graph_capture_s": t_2 - t_1,
                  "run_s": t_3 - t_2}
        total = jobs.max_over_ranks(time.perf_counter() - t_0, device)
        assert len(hist["total_loss"]) == k_steps
        return total, phases, final_host.numel() * 4, runner2.frame_bytes_d2h

    passes = [e2e_pass() for _ in range(5 if full else 3)]
    totals = sorted(p[0] for p in passes)
    e2e_s = totals[len(totals) // 2]  # median pass
    med = min(passes, key=lambda p: abs(p[0] - e2e_s))
    e2e = {"value": info.world_size * k_steps / e2e_s, "unit": "steps/s",
           "h2d_bytes_per_step": (content_h.numel() + style_h.numel()) * 4 / k_steps,
           "d2h_bytes_per_step": 12 + med[2] / k_steps + med[3] / k_steps,
           "steps": k_steps, "phases": med[1], "pass_totals_s": [p[0] for p in passes],
           "how": "pinned host images -> H2D -> set_targets -> OptimizationRunner.run() with "
                  "log_every=1 (the step's losses are read back every step) -> final image D2H; "
                  "MEDIAN of the complete passes"}

    out = {"value": value, "ms_per_step": ms_block / k_steps, "blocks_ms": blocks,
           "steps_per_block": k_steps, "e2e": e2e, "clocks": clocks,
           "gpu_launches_per_step": launches_per_step, "frames_read_back": frames,
           "steps_run_device_resident": steps_done, "height": h, "width": w,
           "workspace_gb": model.engine_for(device).workspace_bytes(h, w) / 1e9}
    if info.rank != 0:
        return out

    # ---- roofline of the dominant kernel (tensor-core conv), measured live with CUDA events ----
    prof = profile_dominant_kernel(model, x, 3)
    peaks = _peaks()
    peak, peak_src = tensor_peak(peaks, clocks)
    achieved = prof["flops"] / (prof["ms"] / 1e3) / 1e12
    traffic = None
    rf = ROOT / "profiles" / "roofline_traffic.json"
    if rf.exists():
        traffic = json.loads(rf.read_text()).get(wl_key)
    step_flops = model.engine_for(device).flops_per_step(h, w)
    out["roofline"] = {
        "bound": "tensor",
        "kernel": "conv_igemm2_tf32_kernel (all tcgen05 conv launches of the step: 3x3 fwd incl. "
                  "pool-fused, 3x3 dgrad incl. un-pooling, first-layer dgrad, 1x1 style-bwd)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "frac_vs_sustained": achieved / (peaks["bf16_sustained"] / 2.0),
        "frac_vs_burst": achieved / (peaks["bf16_burst"] / 2.0),
        "launches_per_step": prof["launches"] // prof["steps"],
        "avg_launch_ms": prof["ms"] / prof["launches"],
        "algorithmic_flops_per_step": prof["flops"] / prof["steps"],
        "share_of_step_time": (prof["ms"] / prof["steps"]) / (ms_block / k_steps),
        "whole_step_tflops": step_flops["total_tri"] / (ms_block / k_steps / 1e3) / 1e12,
        "event_pairs_per_step": prof["runs"] // prof["steps"],
        "measured_on": "3 eager executions of the same step, each queued behind a 40 ms spin kernel "
                       "so that host launch latency is not counted as kernel time; one CUDA-event "
                       "pair on the launching stream around every run of consecutive conv launches "
                       "(runs end at the next loss / first-layer-forward / update kernel)",
    }
    out["roofline_hbm"] = hbm_kernel_rooflines(model, x, device, peaks)
    out["gpu_launches"] = launches_per_step * k_steps
    del runner, model
    torch.cuda.empty_cache()
    return out


def stock_torch_gpu_baseline(device, h: int, w: int, steps: int) -> float:  # noqa: ANN001
    """The reference's algorithm (torchvision VGG19 modules, torch.mm Gram with the 5e5 clamp,
    mse_loss, autograd, torch.optim.Adam, eager launches) through STOCK PyTorch on this GPU: cuDNN
    TF32 convolutions + fp32 cuBLAS -- the "kernel to beat" of SURVEY section 8(d).  Library code,
    reported as a baseline only."""
    import torch
    import torch.nn.functional as F  # noqa: N812

    from style_transfer_visualizer_b200 import synthetic

    feats = synthetic.random_vgg19_features(0).to(device)
    layers = list(feats.children())[:max(STYLE_LAYERS + CONTENT_LAYERS) + 1]
    layers = [torch.nn.ReLU(inplace=False) if isinstance(m, torch.nn.ReLU) else m for m in layers]

    def gram(t):  # noqa: ANN001, ANN202
        b, c, hh, ww = t.shape
        f = t.reshape(b * c, hh * ww)
        return torch.mm(f, f.t()).clamp(max=5e5).div(b * c * hh * ww)

    def taps(img):  # noqa: ANN001, ANN202
        outs, cur = {}, img
        for i, m in enumerate(layers):
            cur = m(cur)
            if i in STYLE_LAYERS or i in CONTENT_LAYERS:
                outs[i] = cur
        return outs

    content = synthetic.synthetic_image(1, h, w).to(device)
    style = synthetic.synthetic_image(2, h, w).to(device)
    with torch.no_grad():
        st = {i: gram(t) for i, t in taps(style).items() if i in STYLE_LAYERS}
        ct = {i: t for i, t in taps(content).items() if i in CONTENT_LAYERS}
    x = content.clone().requires_grad_(True)
    opt = torch.optim.Adam([x], lr=LR)

    def closure():  # noqa: ANN202
        opt.zero_grad()
        o = taps(x)
        sl = torch.stack([F.mse_loss(gram(o[i]), st[i]) for i in STYLE_LAYERS]).sum()
        cl = torch.stack([F.mse_loss(o[i], ct[i]) for i in CONTENT_LAYERS]).sum()
        loss = STYLE_W * sl + CONTENT_W * cl
        loss.backward()
        return loss

    for _ in range(5):
        opt.step(closure)
    torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for _ in range(steps):
        opt.step(closure)
    torch.cuda.synchronize(device)
    value = steps / (time.perf_counter() - t0)
    del opt, x, st, ct, feats, layers
    torch.cuda.empty_cache()
    return value


def measure_jobs64(args, info, device, steps: int, lanes: int) -> dict:  # noqa: ANN001
    """configs[3]: 64 independent jobs; every rank runs its static share through StyleJobPool
    (model, workspaces, pinned buffers and the captured step graph reused across jobs).  The timed
    region is end to end per job: pinned host images in, result image out."""
    import torch

    import style_transfer_visualizer_b200.core_model as cm
    from style_transfer_visualizer_b200 import jobs, synthetic

    wl = WORKLOADS["jobs64"]
    n_jobs = 64
    mine = jobs.partition_jobs(n_jobs, info.world_size, info.rank)
    original = cm.initialize_vgg

    def make_model():  # noqa: ANN202
        cm.initialize_vgg = lambda: synthetic.random_vgg19_features(0)
        try:
            return cm.StyleContentModel(list(STYLE_LAYERS), list(CONTENT_LAYERS))
        finally:
            cm.initialize_vgg = original

    lanes = max(1, lanes)
    pool = jobs.StyleJobPool(make_model, wl["h"], wl["w"], steps=steps, lanes=lanes,
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
    del pool
    torch.cuda.empty_cache()
    return {
        "workload": wl["name"], "value": n_jobs * steps / total_s, "unit": "steps/s",
        "jobs": n_jobs, "jobs_per_gpu": len(mine), "concurrent_lanes_per_gpu": lanes,
        "steps_per_job": steps, "jobs_per_second": n_jobs / total_s, "scaling": "strong",
        "ms_per_step": total_s / (len(mine) * steps) * 1e3,
        "timing": "host wall clock per rank over its jobs (H2D of both images and D2H of the "
                  "result inside), max over ranks",
        "e2e": {"value": n_jobs * steps / total_s, "unit": "steps/s",
                "h2d_bytes_per_step": 2 * 3 * wl["h"] * wl["w"] * 4 / steps,
                "d2h_bytes_per_step": 3 * wl["h"] * wl["w"] * 4 / steps},
        "final_loss_first_job": losses[0],
    }


def _golden_4k_parity(model) -> dict | None:  # noqa: ANN001
    """Sharded first closure against tests/golden/adam_random_4k_c5.npz (outputs of the unmodified
    reference on the CPU): per-layer losses and the gathered input gradient (stored ::8 samples).
    Collective: every rank calls it; only rank 0 gets the numbers."""
    import numpy as np
    import torch

    from style_transfer_visualizer_b200 import synthetic

    path = ROOT / "tests" / "golden" / "adam_random_4k_c5.npz"
    if not path.exists():
        return None
    gold = np.load(path, allow_pickle=False)
    content = synthetic.synthetic_image(1, 2160, 3840)
    start = torch.randn(content.shape, generator=torch.Generator().manual_seed(3))
    x = model.band_of(start).requires_grad_(True)
    sl, cl = model(x)
    (STYLE_W * torch.stack(sl).sum() + CONTENT_W * torch.stack(cl).sum()).backward()
    full = model.gather_image(x.grad)
    if full is None:
        return None
    g = full.detach().cpu().numpy()[..., ::8, ::8].astype(np.float64)
    ref = gold["first_grad"].astype(np.float64)
    ls = np.array([float(v.detach()) for v in sl])
    lc = np.array([float(v.detach()) for v in cl])
    return {
        "fixture": "tests/golden/adam_random_4k_c5.npz (unmodified reference, CPU fp32)",
        "grad_rel_l2": float(np.linalg.norm(g - ref) / np.linalg.norm(ref)),
        "layer_style_rel_max": float(np.max(np.abs(ls - gold["layer_style"]) / gold["layer_style"])),
        "content_rel": float(np.max(np.abs(lc - gold["layer_content"]) / gold["layer_content"])),
    }


def measure_sharded_4k(args, info, device, steps: int, sampler) -> dict:  # noqa: ANN001
    """configs[4]: ONE 3840x2160 image split into row bands over all ranks (strong scaling)."""
    import torch

    from style_transfer_visualizer_b200 import _native as nat
    from style_transfer_visualizer_b200 import synthetic
    from style_transfer_visualizer_b200.sharded import ShardedFusedStep, ShardedStyleContentModel

    wl = WORKLOADS["4k"]
    h, w = wl["h"], wl["w"]
    model = ShardedStyleContentModel(synthetic.random_vgg19_features(0), list(STYLE_LAYERS),
                                     list(CONTENT_LAYERS), device)
    content = synthetic.synthetic_image(1, h, w)
    style = synthetic.synthetic_image(2, h, w)
    model.set_targets(style, content)
    parity = _golden_4k_parity(model)
    x = model.band_of(content)
    fused = ShardedFusedStep(model, x, lr=LR, style_w=STYLE_W, content_w=CONTENT_W)
    launches0 = nat.launch_count()
    for _ in range(3):
        fused.step()
    per_step = (nat.launch_count() - launches0) // 3
    graph_mode = "eager"
    if args.sharded_graph:
        fused.capture()
        graph_mode = "cuda-graph (halo exchange + all-reduce captured)"
    blocks, t0, t1 = time_blocks(fused.step, steps, device, min_blocks=3, min_total_s=0.5,
                                 warm_s=0.3)
    ms_block = statistics.median(blocks)
    loss = float(fused.scores[2])
    out = {
        "workload": wl["name"], "value": steps / (ms_block / 1e3), "unit": "steps/s",
        "n_gpus": info.world_size, "ms_per_step": ms_block / steps, "blocks_ms": blocks,
        "steps_per_block": steps, "scaling": "strong",
        "parallelism": f"row bands x{info.world_size} (16-row aligned)",
        "halo_exchange": getattr(model.engine, "halo_mode", "nccl send/recv"),
        "step_launch": graph_mode, "gpu_launches_per_step": per_step, "final_loss": loss,
        "parity_vs_golden": parity,
        "clocks": sampler.window(t0, t1) if sampler else None,
        "whole_job_tflops": model.engine.base.flops_per_step(h, w)["total_tri"]
        / (ms_block / steps / 1e3) / 1e12,
    }
    torch.cuda.synchronize(device)
    fused.graph = None
    del fused, model
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return out


def measure_4k_single_gpu(device, steps: int) -> float:  # noqa: ANN001
    """steps/s of the same 3840x2160 step on ONE GPU with the unsharded engine (rank 0 only): the
    denominator of the sharded path's strong-scaling efficiency."""
    import torch

    from style_transfer_visualizer_b200.core_model import initialize_input
    from style_transfer_visualizer_b200.fused_step import FusedStep
    from style_transfer_visualizer_b200.optim import FusedAdam

    model, content_h, style_h = build_job(device, 2160, 3840, 0)
    content = content_h.to(device)
    model.set_targets(style_h.to(device), content)
    x = initialize_input(content, "content")
    fused = FusedStep.try_create(model, x, FusedAdam([x], lr=LR), STYLE_W, CONTENT_W)
    for _ in range(3):
        fused.step()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fused.step()
    e1.record()
    torch.cuda.synchronize(device)
    value = steps / (e0.elapsed_time(e1) / 1e3)
    del fused, model
    torch.cuda.empty_cache()
    return value


# ------------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------------
def _init(args):  # noqa: ANN001, ANN202, ARG001
    import torch

    from style_transfer_visualizer_b200 import _native as nat
    from style_transfer_visualizer_b200 import jobs

    os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
    info = jobs.init_distributed()
    device = torch.device("cuda", info.local_rank)
    torch.cuda.set_device(device)
    nat.require_device(device)  # fails loudly if the .so is missing or the GPU is not sm_100
    sampler = None
    if info.rank == 0:
        sampler = ClockSampler(info.local_rank)
        sampler.start()
    return info, device, sampler


def _finish(sampler, used_captured_nccl: bool) -> None:  # noqa: ANN001, FBT001
    from style_transfer_visualizer_b200 import jobs

    if sampler:
        sampler.stop()
    jobs.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    if used_captured_nccl:
        # leave without destroy_process_group(): destroying an NCCL communicator that was used
        # inside a captured graph blocked at exit on this stack (torch 2.11 / NCCL 2.28)
        os._exit(0)
    jobs.shutdown()


def run_gpu_arm(args, wl_key: str) -> None:  # noqa: ANN001
    info, device, sampler = _init(args)
    wl = WORKLOADS[wl_key]
    main = measure_single_image(wl_key, args, info, device, sampler, full=True)
    extras: dict = {}
    captured_nccl = False
    if not args.no_extras:
        if info.world_size == 1:
            other = "1080p" if wl_key == "512" else "512"
            sub = argparse.Namespace(**{**vars(args), "steps": min(args.steps, 20)})
            extras[other] = measure_single_image(other, sub, info, device, sampler, full=False)
            extras[other]["workload"] = WORKLOADS[other]["name"]
            extras["jobs64"] = measure_jobs64(args, info, device, 30, args.lanes)
            lib = {}
            for key in ("512", "1080p"):
                hh, ww = WORKLOADS[key]["h"], WORKLOADS[key]["w"]
                lib[key] = stock_torch_gpu_baseline(device, hh, ww, 40 if key == "512" else 10)
            extras["gpu_library_baseline"] = {
                "steps_per_s": lib, "unit": "steps/s",
                "what": "the reference's algorithm through stock PyTorch on this GPU (cuDNN TF32 "
                        "convs, fp32 cuBLAS mm, eager autograd, torch.optim.Adam): the library "
                        "'kernel to beat' of SURVEY section 8(d), not the CPU baseline"}
        else:
            sh = measure_sharded_4k(args, info, device, 10, sampler)
            captured_nccl = bool(args.sharded_graph)
            if info.rank == 0:
                n1 = measure_4k_single_gpu(device, 10)
                sh["n1_value_unsharded"] = n1
                sh["efficiency_vs_n1"] = sh["value"] / (info.world_size * n1)
            extras["sharded_4k"] = sh
    if info.rank == 0:
        cpu = None
        if info.world_size == 1 and not args.no_cpu_baseline:
            v, done, threads = cpu_reference_run(wl["h"], wl["w"], budget_s=12.0, max_steps=20)
            cpu = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port",
                   "sample": f"{done} Adam steps at {wl['w']}x{wl['h']} (12 s budget) of the oracle "
                             "port of the reference's path, all host threads"}
        line = {
            "metric": METRIC, "value": main["value"], "unit": "steps/s",
            "n_gpus": info.world_size, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": wl["name"], "height": wl["h"], "width": wl["w"],
                       "optimizer": "adam", "lr": LR, "style_w": STYLE_W, "content_w": CONTENT_W,
                       "jobs_per_gpu": 1, "parallelism": f"independent jobs x{info.world_size}",
                       "l2": "no flush: each step streams ~%.1f GB of activations/gradients, far "
                             "above the 126 MB L2" % main["workspace_gb"],
                       "storage": "fp32 activations, TF32 multiply / FP32 accumulate",
                       "timing": "median of %d blocks of %d steps (CUDA events, barrier + "
                                 "synchronize around every block, max over ranks) after %d warm-up "
                                 "steps and a >= 0.5 s warm run"
                                 % (len(main["blocks_ms"]), args.steps, max(3, args.warmup))},
            "blocks_ms": main["blocks_ms"],
            "e2e": main["e2e"], "gpu_launches": main.get("gpu_launches"),
            "gpu_launches_per_step": main["gpu_launches_per_step"],
            "clocks": main["clocks"], "roofline": main.get("roofline"),
            "roofline_hbm": main.get("roofline_hbm"), "cpu_baseline": cpu,
            "frames_read_back": main["frames_read_back"],
            "workloads": extras,
        }
        print(json.dumps(line), flush=True)
    _finish(sampler, captured_nccl)


def run_jobs_arm(args) -> None:  # noqa: ANN001
    info, device, sampler = _init(args)
    res = measure_jobs64(args, info, device, args.steps, args.lanes)
    if info.rank == 0:
        line = {
            "metric": METRIC, "value": res["value"], "unit": "steps/s", "n_gpus": info.world_size,
            "steps": args.steps, "warmup": 1, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32",
            "data": "synthetic",
            "config": {"workload": res["workload"], "jobs": res["jobs"],
                       "jobs_per_gpu": res["jobs_per_gpu"],
                       "concurrent_lanes_per_gpu": res["concurrent_lanes_per_gpu"],
                       "steps_per_job": res["steps_per_job"], "height": 512, "width": 512,
                       "timing": res["timing"]},
            "jobs_per_second": res["jobs_per_second"], "e2e": res["e2e"],
            "final_loss_first_job": res["final_loss_first_job"],
        }
        print(json.dumps(line), flush=True)
    _finish(sampler, False)


def run_sharded_arm(args) -> None:  # noqa: ANN001
    import torch

    info, device, sampler = _init(args)
    if info.world_size == 1 and not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        torch.distributed.init_process_group("nccl", rank=0, world_size=1,
                                             device_id=torch.device("cuda", 0))
    res = measure_sharded_4k(args, info, device, args.steps, sampler)
    if info.rank == 0:
        wl = WORKLOADS["4k"]
        line = {
            "metric": METRIC, "value": res["value"], "unit": "steps/s", "n_gpus": info.world_size,
            "steps": args.steps, "warmup": 3, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32",
            "data": "synthetic", "blocks_ms": res["blocks_ms"],
            "config": {"workload": wl["name"], "height": wl["h"], "width": wl["w"],
                       "optimizer": "adam", "parallelism": res["parallelism"],
                       "halo_exchange": res["halo_exchange"],
                       "l2": "no flush: working set far above the 126 MB L2",
                       "step_launch": res["step_launch"]},
            "e2e": None, "gpu_launches": res["gpu_launches_per_step"] * args.steps,
            "gpu_launches_per_step": res["gpu_launches_per_step"], "clocks": res["clocks"],
            "final_loss": res["final_loss"], "parity_vs_golden": res["parity_vs_golden"],
            "whole_job_tflops": res["whole_job_tflops"],
        }
        print(json.dumps(line), flush=True)
    _finish(sampler, bool(args.sharded_graph))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="512", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the other BASELINE configs measured into the same JSON line")
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
        args.steps = wl["steps"] if args.impl == "b200" else 20
    if args.impl == "reference":
        run_reference_arm(args, wl)
    elif args.workload == "4k":
        run_sharded_arm(args)
    elif args.workload == "jobs64":
        run_jobs_arm(args)
    else:
        run_gpu_arm(args, args.workload)


if __name__ == "__main__":
    main()
