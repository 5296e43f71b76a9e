"""Bottleneck experiments on single conv layers: STV_CONV_DEBUG bits disable weight loads (1),
activation loads (2), epilogue stores (4), MMAs (8); results are garbage, only the time matters.
Needs the experiments build of the library: `python build_native.py --force --experiments` (the
product build does not read STV_CONV_DEBUG); rebuild without the flag afterwards.
"""
import os, sys, subprocess
code = r'''
import sys, os, torch
sys.path.insert(0, ".")
from style_transfer_visualizer_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
for (h, w, c, n, bn, mh, pair) in eval(os.environ["CFGS"]):
    x = torch.randn(h, w, c, device=dev, generator=g)
    wt = torch.randn(n, c, 3, 3, device=dev, generator=g) * 0.05
    wf, _ = ops.pack_conv_weights(wt)
    bias = torch.randn(n, device=dev, generator=g)
    post = torch.empty(h, w, n, device=dev)
    ops.conv_set_tuning(pair)
    f = lambda: ops.conv_igemm2_ex(x, wf, taps=9, bias=bias, out_post=post, block_n=bn, m_halves=mh)
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"  {h}x{w} C{c} N{n} bn{bn} mh{mh} pair{pair}: {ms*1e3:7.1f} us {2.0*h*w*c*n*9/ms/1e9:7.1f} TF/s", flush=True)
'''
cfgs = eval(sys.argv[1]) if len(sys.argv) > 1 else [(64, 64, 512, 512, 128, 1, 0), (64, 64, 512, 512, 128, 1, 1), (64, 64, 512, 512, 256, 1, 1),
                                                    (128, 128, 256, 256, 256, 1, 0), (128, 128, 256, 256, 256, 1, 1), (32, 32, 512, 512, 64, 1, 0)]
for dbg in (0, 1, 2, 3, 4, 7, 8, 12, 15):
    print(f"STV_CONV_DEBUG={dbg} (1 skip B loads, 2 skip A loads, 4 skip stores, 8 skip MMAs)", flush=True)
    env = dict(os.environ, STV_CONV_DEBUG=str(dbg), CFGS=repr(cfgs))
    subprocess.run([sys.executable, "-c", code], env=env, timeout=120, check=False)
