"""Handoff-chain experiment: one 64x64 512->512 conv under STV_CONV_DEBUG modes (15 = control skeleton
only, 7 = MMAs only, 3 = no loads, 0 = full) for several ring configurations (profiles/r1_conv_chain_latency.log).
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
h, w, c, n = 64, 64, 512, 512
x = torch.randn(h, w, c, device=dev, generator=g)
wt = torch.randn(n, c, 3, 3, device=dev, generator=g) * 0.05
wf, _ = ops.pack_conv_weights(wt)
bias = torch.randn(n, device=dev, generator=g)
post = torch.empty(h, w, n, device=dev)
out = []
for (bn, mh, pair, as_, bs, tps) in eval(os.environ["CFGS"]):
    ops.conv_set_tuning(pair, as_, bs, tps)
    f = lambda: ops.conv_igemm2_ex(x, wf, taps=9, bias=bias, out_post=post, block_n=bn, m_halves=mh)
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(10): f()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    out.append(f"{best*1e3:6.1f}")
print("   ".join(out), flush=True)
'''
cfgs = [(128,1,0,2,3,3), (128,1,0,3,3,3), (128,1,0,4,2,3), (128,1,0,2,2,3), (128,1,1,2,3,3), (128,1,1,3,3,3), (128,1,1,4,4,3), (128,1,1,5,4,3), (128,1,1,2,6,1), (128,1,0,3,4,1), (256,1,1,3,3,3), (64,1,0,2,2,3)]
print("cfgs (bn,mh,pair,as,bs,tps):", cfgs)
for dbg in (15, 7, 3, 0):
    env = dict(os.environ, STV_CONV_DEBUG=str(dbg), CFGS=repr(cfgs))
    print(f"dbg={dbg:2d}: ", end="", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env, timeout=120, check=False)
