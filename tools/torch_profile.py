"""torch.profiler breakdown of bench steps (GPU kernel time by name + CPU time)."""
import sys, math, torch
sys.path.insert(0, ".")
import bench
from gaustudio_b200 import _C
from gaustudio_b200.synthetic import build_config
from gaustudio_b200.camera import orbit_cameras
impl = sys.argv[1] if len(sys.argv) > 1 else "new"
model, _, c = build_config("cfg3", K=1)
dev = torch.device("cuda")
cams = orbit_cameras(200, c["radius"], c["elev"], c["W"], c["H"], c["fovx"], c["fovy"], indices=list(range(30)))
hc = [bench.HostCamera(cm).upload(dev) for cm in cams]
model.to(dev).requires_grad_(True)
step = bench.make_step(impl, model, dev, c["H"], c["W"], fused=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
if impl == "new": _C.set_pipelined(True)
for i in range(8): step(hc[i])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(8, 28): step(hc[i])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
import time
t=time.time()
for i in range(8, 28): step(hc[i])
t1=time.time()-t; torch.cuda.synchronize(); t2=time.time()-t
print(f"host-only time per step {t1/20*1e3:.3f} ms ; with sync {t2/20*1e3:.3f} ms")
