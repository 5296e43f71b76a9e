mkdir -p gpurun_out
timeout 600 python tools/gpu_selftest.py --case conv_compact > gpurun_out/r2_selftest_cc_15.log 2>&1; echo "rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_cc_15.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_cc_15.log)"
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for (h, w) in [(1080, 1920), (512, 512), (2160, 3840), (270, 3840)]:
    img = torch.rand(1, 3, h, w, device=dev, generator=g)
    wt = torch.randn(64, 3, 3, 3, device=dev, generator=g) * 0.2
    b = torch.randn(64, device=dev, generator=g) * 0.1
    pre = torch.empty(h, w, 64, device=dev); post = torch.empty(h, w, 64, device=dev)
    bits = ops.relu_bits_buffer(h, w, 64, dev)
    f = lambda: ops.conv3x3_first_fwd(img, wt, b, pre, post, round_pre=True, out_bits=bits)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    e.record(); torch.cuda.synchronize()
    us = a.elapsed_time(e) * 1e3 / 20
    byts = h * w * 64 * 4 * 2 + h * w * 8 + h * w * 12
    print(f"conv1_1 fwd {h}x{w}: {us:7.1f} us  {byts / us / 1e6:6.2f} TB/s", flush=True)
PY
