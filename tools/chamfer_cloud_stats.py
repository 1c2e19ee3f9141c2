import os, sys
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT,"tools")); sys.path.insert(0, os.path.join(ROOT,"temporal-pointcloud-upsampling-gan_b200"))
import torch, refstep
from tpugan_b200.recording import log
ctx = refstep.build("fluid", B=8, n_lo=2048, ratio=4, backend="cuda")
log.start(capture=True)
refstep.step(ctx, 12)
calls = log.stop()
for c in calls:
    if c.op == "chamfer":
        for k in ("src","tgt"):
            x = c.inputs[k]
            print(k, tuple(x.shape), "min", x.amin((0,1)).tolist(), "max", x.amax((0,1)).tolist(), "std", x.std((0,1)).tolist())
            q = torch.quantile(x[0,:,0], torch.tensor([0.001,0.01,0.5,0.99,0.999], device=x.device)); print("  x quantiles cloud0", q.tolist())
