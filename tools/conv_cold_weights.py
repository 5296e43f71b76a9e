"""How much does a conv launch lose when its packed weights come from HBM instead of L2?
Per layer: (a) hot back-to-back loop; (b) L2 flushed, then the activations touched (weights cold,
as inside the optimisation step where every layer's weights were last used a whole step ago);
(c) L2 flushed, activations AND weights touched (method check: should equal a single hot launch)."""
import sys, torch
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
junk = torch.empty(96 * 1024 * 1024, device=dev)   # 384 MB > 126 MB L2
layers = [(512, 512, 64, 64), (256, 256, 128, 128), (128, 128, 256, 256), (64, 64, 512, 512), (32, 32, 512, 512),
          (270, 480, 256, 256), (135, 240, 512, 512)]
tot = [0.0, 0.0, 0.0]
for h, w, c, n in layers:
    x = torch.randn(h, w, c, device=dev, generator=g)
    wt = torch.randn(n, c, 3, 3, device=dev, generator=g) * 0.05
    wf, _ = ops.pack_conv_weights(wt)
    bias = torch.randn(n, device=dev, generator=g)
    out = torch.empty(h, w, n, device=dev)
    f = lambda: ops.conv3x3_fwd(x, wf, bias, None, out)
    for _ in range(3): f()
    res = []
    for mode in ("hot", "cold_w", "warm_all"):
        ts = []
        for _ in range(6):
            if mode != "hot":
                junk.fill_(1.0)
                x.sum(); out.sum()
                if mode == "warm_all":
                    wf.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(sorted(ts)[len(ts) // 2])
    for i in range(3): tot[i] += res[i]
    print(f"{h}x{w} {c}->{n}: single hot launch {res[0]:7.1f} us | weights cold {res[1]:7.1f} us | flushed+all touched {res[2]:7.1f} us  (weights {wf.numel()*4/1e6:.1f} MB)", flush=True)
print(f"sum: {tot[0]:.1f} / {tot[1]:.1f} / {tot[2]:.1f}")
