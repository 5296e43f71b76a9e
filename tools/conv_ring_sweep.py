"""Ring depth / CTA-pair sweep on representative layers (profiles/r1_ring_sweep.log)."""
import sys, torch
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def run(h, w, c, n, bn, mh, pair, as_, bs, tps, heavy=0):
    x = torch.randn(h, w, c, device=dev, generator=g)
    wt = torch.randn(n, c, 3, 3, device=dev, generator=g) * 0.05
    wf, _ = ops.pack_conv_weights(wt)
    bias = torch.randn(n, device=dev, generator=g)
    out = torch.empty(h, w, n, device=dev)
    kw = dict(taps=9, block_n=bn, m_halves=mh)
    if heavy:
        mask = torch.randn(h, w, n, device=dev, generator=g); add = torch.randn(h, w, n, device=dev, generator=g)
        kw.update(mask_src=mask, add_src=add, out_pre=out)
    else:
        kw.update(bias=bias, out_post=out)
    ops.conv_set_tuning(pair, as_, bs, tps)
    f = lambda: ops.conv_igemm2_ex(x, wf, **kw)
    try:
        for _ in range(3): f()
        torch.cuda.synchronize()
    except Exception as e:
        print(f"  {h}x{w} C{c}->N{n} bn{bn} mh{mh} pair{pair} as{as_} bs{bs} tps{tps} heavy{heavy}: FAIL {str(e)[:70]}"); return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(8): f()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 8)
    print(f"  {h}x{w} C{c}->N{n} bn{bn} mh{mh} pair{pair} as{as_} bs{bs} tps{tps} heavy{heavy}: {best*1e3:7.1f} us {2.0*h*w*c*n*9/best/1e9:6.1f} TF/s", flush=True)

for heavy in (0, 1):
    print("64x64 512->512")
    run(64, 64, 512, 512, 128, 1, 0, 0, 0, 0, heavy)
    for (as_, bs) in [(2, 3), (3, 3), (4, 4), (5, 5), (4, 5), (5, 4), (6, 4)]:
        run(64, 64, 512, 512, 128, 1, 1, as_, bs, 3, heavy)
    for (as_, bs) in [(3, 3), (4, 2), (3, 9), (4, 8)]:
        run(64, 64, 512, 512, 128, 1, 0, as_, bs, 3 if bs < 6 else 1, heavy)
    print("128x128 256->256")
    run(128, 128, 256, 256, 256, 1, 0, 0, 0, 0, heavy)
    for (as_, bs) in [(2, 3), (3, 3), (4, 3), (3, 2), (4, 2)]:
        run(128, 128, 256, 256, 256, 1, 1, as_, bs, 3, heavy)
    for (as_, bs) in [(4, 5), (5, 5)]:
        run(128, 128, 256, 256, 128, 1, 1, as_, bs, 3, heavy)
    print("256x256 128->128")
    run(256, 256, 128, 128, 128, 2, 0, 0, 0, 0, heavy)
    for (mh, as_, bs) in [(1, 4, 5), (1, 5, 5), (2, 3, 4), (2, 3, 3), (2, 2, 3)]:
        run(256, 256, 128, 128, 128, mh, 1, as_, bs, 3, heavy)
    print("270x480 256->256")
    run(270, 480, 256, 256, 256, 1, 0, 0, 0, 0, heavy)
    for (as_, bs) in [(2, 3), (3, 3), (4, 3)]:
        run(270, 480, 256, 256, 256, 1, 1, as_, bs, 3, heavy)
    print("135x240 512->512")
    run(135, 240, 512, 512, 256, 2, 0, 0, 0, 0, heavy)
    for (mh, as_, bs, tps) in [(2, 2, 6, 1), (1, 3, 3, 3), (1, 4, 3, 3), (2, 2, 2, 3), (2, 3, 6, 1)]:
        run(135, 240, 512, 512, 256, mh, 1, as_, bs, tps, heavy)
ops.conv_set_tuning()
