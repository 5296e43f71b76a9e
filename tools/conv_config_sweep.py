"""Sweep conv tile configurations per VGG layer shape (fwd and dgrad-style heavy epilogue)."""
import sys, torch
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "1080p"
if which == "1080p":
    layers = [(1080,1920,64,64), (540,960,64,128), (540,960,128,64), (540,960,128,128), (270,480,128,256), (270,480,256,128),
              (270,480,256,256), (135,240,256,512), (135,240,512,256), (135,240,512,512), (67,120,512,512)]
else:
    layers = [(512,512,64,64), (256,256,64,128), (256,256,128,64), (256,256,128,128), (128,128,128,256), (128,128,256,128),
              (128,128,256,256), (64,64,256,512), (64,64,512,256), (64,64,512,512), (32,32,512,512)]
# (bn, mh, pair, as, bs, tps)
def cands(n):
    out = [(0, 0, 0, 0, 0, 0)]  # current rule table
    for bn in (256, 128, 64):
        if n % bn: continue
        for mh in (1, 2):
            for tps in (1, 3):
                out.append((bn, mh, 1, 0, 0, tps))
            if bn <= 128:
                out.append((bn, mh, 0, 0, 0, 0))
                out.append((bn, mh, 0, 3, 3, 0))
                out.append((bn, mh, 0, 4, 4, 0))
    return out
for (h, w, c, n) in layers:
    x = torch.randn(h, w, c, device=dev, generator=g)
    wt = torch.randn(n, c, 3, 3, device=dev, generator=g) * 0.05
    wf, _ = ops.pack_conv_weights(wt)
    bias = torch.randn(n, device=dev, generator=g)
    post = torch.empty(h, w, n, device=dev)
    mask = torch.randn(h, w, n, device=dev, generator=g)
    add = torch.randn(h, w, n, device=dev, generator=g)
    for heavy in (0, 1):
        rows = []
        for (bn, mh, pair, as_, bs, tps) in cands(n):
            ops.conv_set_tuning(pair if bn else -1, as_, bs, tps)
            kw = dict(taps=9, block_n=bn, m_halves=mh)
            if heavy: kw.update(mask_src=mask, add_src=add, out_pre=post)
            else: kw.update(bias=bias, out_post=post)
            f = lambda: ops.conv_igemm2_ex(x, wf, **kw)
            try:
                for _ in range(2): f()
                torch.cuda.synchronize()
            except Exception as e:
                rows.append((1e9, f"bn{bn} mh{mh} pair{pair} tps{tps}: fail {str(e)[:60]}")); continue
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(8): f()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 8
            rows.append((ms, f"bn{bn} mh{mh} pair{pair} tps{tps} as{as_}: {ms*1e3:7.1f} us {2.0*h*w*c*n*9/ms/1e9:6.1f} TF/s"))
        base = rows[0][0]
        rows_sorted = sorted(rows[1:])[:4]
        print(f"{h}x{w} C{c}->N{n} heavy={heavy}: rule {rows[0][1]}", flush=True)
        for ms, txt in rows_sorted:
            print(f"      {txt}  ({base/ms:4.2f}x)", flush=True)
ops.conv_set_tuning()
