for mh in 1 2; do for tw in 8 16 32; do STV_FD_MH=$mh STV_FD_TW=$tw python - <<PY
import torch,sys
sys.path.insert(0,'.')
from style_transfer_visualizer_b200 import ops
dev=torch.device('cuda')
pre=torch.randn(1080,1920,64,device=dev); w1=torch.randn(64,3,3,3,device=dev); w16=ops.pack_first_dgrad_weights(w1); dimg=torch.empty(1,3,1080,1920,device=dev)
for _ in range(3): ops.conv3x3_first_dgrad_tc(pre,w16,dimg)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(10): ops.conv3x3_first_dgrad_tc(pre,w16,dimg)
e1.record(); torch.cuda.synchronize(); print("first_dgrad_tc mh=$mh tw=$tw", e0.elapsed_time(e1)/10, "ms")
PY
done; done
