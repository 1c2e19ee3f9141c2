import os, sys
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT,"tools")); sys.path.insert(0, os.path.join(ROOT,"temporal-pointcloud-upsampling-gan_b200"))
import torch, refstep
from torch.profiler import ProfilerActivity, profile
ctx = refstep.build("fluid", B=8, n_lo=2048, ratio=4, backend="cuda", capturable=True)
gs = refstep.graphed_step(ctx, capture=True)
for n in (12,14,16): gs.step(n)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs.step(18); torch.cuda.synchronize()
t = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=80)
open(os.path.join(ROOT,"gpurun_out","graphed_prof.txt"),"w").write(t)
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev); n = sum(e.count for e in ev)
print("total cuda ms", tot/1e3, "kernels", n)
