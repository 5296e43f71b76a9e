"""Per-layer conv times of one optimisation step (rule-table configs), fwd + dgrad shapes with their
multiplicities; prints a weighted total.  Usage: conv_layer_times.py [512|1080p]"""
import sys, torch
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "512"
H, W = (512, 512) if which == "512" else (1080, 1920)
# (scale, C, N, count_fwd, count_dgrad)   dgrad runs the transposed problem N->C with the ReLU gate
stack = [(1, 64, 64, 1, 1), (2, 64, 128, 1, 1), (2, 128, 128, 1, 1), (4, 128, 256, 1, 1), (4, 256, 256, 3, 3),
         (8, 256, 512, 1, 1), (8, 512, 512, 3, 3), (16, 512, 512, 1, 1)]
total = 0.0
for sc, c, n, nf, nd in stack:
    h, w = H // sc, W // sc
    for heavy, cin, cout, cnt in ((0, c, n, nf), (1, n, c, nd)):
        x = torch.randn(h, w, cin, device=dev, generator=g)
        wt = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
        wf, _ = ops.pack_conv_weights(wt)
        bias = torch.randn(cout, device=dev, generator=g)
        out = torch.empty(h, w, cout, device=dev)
        mask = torch.randn(h, w, cout, device=dev, generator=g)
        add = torch.randn(h, w, cout, device=dev, generator=g)
        kw = dict(taps=9)
        if heavy: kw.update(mask_src=mask, add_src=add, out_pre=out)
        else: kw.update(bias=bias, out_post=out)
        f = lambda: ops.conv_igemm2_ex(x, wf, **kw)
        for _ in range(3): f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            a.record()
            for _ in range(6): f()
            b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 6)
        total += best * cnt
        print(f"{h}x{w} {cin}->{cout} {'dgrad' if heavy else 'fwd  '} x{cnt}: {best*1e3:7.1f} us {2.0*h*w*cin*cout*9/best/1e9:6.1f} TF/s", flush=True)
print(f"TOTAL {which}: {total*1e3:.1f} us")
