mkdir -p gpurun_out
run() { for sz in 512 1080p; do timeout 300 python tools/plan_sweep.py --size $sz --ab-resident --reps $([ $sz = 512 ] && echo 200 || echo 40) 2>&1 | grep "round 1 rule" | sed "s/^/$1 $sz: /"; done
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for (hw, c) in [(1080*1920, 64), (540*960, 128), (270*480, 256), (135*240, 512), (512*512, 64), (64*64, 512)]:
    x = torch.randn(hw, c, device=dev, generator=g)
    ws = ops.gram_workspace(hw, c, dev)
    tgt = torch.zeros(c, c, device=dev); s_out = torch.empty(c, c, device=dev); loss = torch.zeros(1, device=dev)
    f = lambda: ops.gram_loss_fwd(x, ws, target=tgt, s_out=s_out, loss_out=loss)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    e.record(); torch.cuda.synchronize()
    us = a.elapsed_time(e) * 1e3 / 20
    print(f"   gram fwd hw={hw} C={c}: {us:7.1f} us  {hw * c * 4 / us / 1e6:5.2f} TB/s", flush=True)
PY
}
timeout 600 python tools/gpu_selftest.py --case gram > gpurun_out/r2_selftest_gram_24.log 2>&1; echo "gram rc=$? pass=$(grep -c ^PASS gpurun_out/r2_selftest_gram_24.log) fail=$(grep -c ^FAIL gpurun_out/r2_selftest_gram_24.log)"
run "gram 64 px/stage"
STV_NVCC_EXTRA="-DSTV_GRAM_PIX=32" python build_native.py --force > /dev/null 2>&1; echo "build rc=$?"
run "gram 32 px/stage"
