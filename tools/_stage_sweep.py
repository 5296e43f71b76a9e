import os, sys, torch
sys.path.insert(0, '.')
from style_transfer_visualizer_b200 import ops
dev = torch.device('cuda')
g = torch.Generator(device='cuda').manual_seed(0)
shapes = [(1080,1920,64,64),(540,960,128,128),(540,960,64,128),(270,480,256,256),(135,240,512,512),(512,512,64,64),(256,256,128,128),(128,128,256,256),(64,64,512,512),(32,32,512,512)]
out = []
for (h,w,cin,cout) in shapes:
    x = torch.randn(h,w,cin,device=dev,generator=g); wt = torch.randn(cout,cin,3,3,device=dev,generator=g)*0.05
    b = torch.zeros(cout,device=dev); wf,wd = ops.pack_conv_weights(wt); post = torch.empty(h,w,cout,device=dev)
    for _ in range(3): ops.conv3x3_fwd(x,wf,b,None,post)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(10): ops.conv3x3_fwd(x,wf,b,None,post)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
    out.append(f"{2.0*9*cin*cout*h*w/ms/1e9:.0f}")
print("STAGES", os.environ.get("STV_CONV_STAGES","default"), " ".join(out))
