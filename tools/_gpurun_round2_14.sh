mkdir -p gpurun_out
for c in elementwise conv conv_compact; do timeout 600 python tools/gpu_selftest.py --case $c > gpurun_out/r2_selftest_${c}_14.log 2>&1; echo "$c rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_${c}_14.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_${c}_14.log)"; grep ^FAIL gpurun_out/r2_selftest_${c}_14.log | head -5; done
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for (h, w) in [(1080, 1920), (512, 512), (2160, 3840)]:
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
    ref = torch.nn.functional.conv2d(img.double(), wt.double(), b.double(), padding=1)[0].permute(1, 2, 0)
    err = float((pre.double() - ref).abs().max())
    print(f"   max abs err vs fp64 conv: {err:.3e} (tf32 rounding of pre: <= {float(ref.abs().max()) * 2**-11:.3e})")
PY
python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2_bench512_v14.json 2> gpurun_out/r2_bench512_v14.err; echo "bench rc=$?"
python bench.py --no-extras --no-cpu-baseline --workload 1080p > gpurun_out/r2_bench1080_v14.json 2> gpurun_out/r2_bench1080_v14.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("r2_bench512_v14", "r2_bench1080_v14"):
    for ln in reversed(open(f"gpurun_out/{f}.json").read().strip().splitlines()):
        if ln.startswith("{"):
            d = json.loads(ln); print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["achieved"],1)); 
            for rr in d.get("roofline_hbm") or []: print("   hbm", rr["kernel"], round(rr["achieved"]), round(rr["frac"],2))
            break
PY
