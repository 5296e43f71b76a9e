mkdir -p gpurun_out
timeout 600 python tools/gpu_selftest.py --case conv2 > gpurun_out/r2_selftest_conv2_16.log 2>&1; echo "conv2 rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_conv2_16.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_conv2_16.log)"; grep "first dgrad\|^FAIL\|watchdog\|rror" gpurun_out/r2_selftest_conv2_16.log | head -40
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for (h, w) in [(1080, 1920), (512, 512), (2160, 3840), (272, 3840)]:
    dy = torch.randn(h, w, 64, device=dev, generator=g)
    wt = torch.randn(64, 3, 3, 3, device=dev, generator=g) * 0.2
    dimg = torch.empty(1, 3, h, w, device=dev)
    w16 = ops.pack_first_dgrad_weights(wt); wr = ops.pack_first_dgrad_rows(wt)
    for name, f in (("N=16 igemm", lambda: ops.conv3x3_first_dgrad_tc(dy, w16, dimg)), ("x taps in N", lambda: ops.conv3x3_first_dgrad_rows(dy, wr, dimg))):
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): f()
        e.record(); torch.cuda.synchronize()
        us = a.elapsed_time(e) * 1e3 / 20
        print(f"first dgrad {h}x{w} {name:12s}: {us:7.1f} us  {(h*w*64*4 + h*w*12) / us / 1e6:6.2f} TB/s", flush=True)
PY
