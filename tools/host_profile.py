"""cProfile of the host side of bench steps (no sync inside the loop)."""
import sys, cProfile, pstats, time, torch
sys.path.insert(0, ".")
import bench
from gaustudio_b200 import _C
from gaustudio_b200.synthetic import build_config
from gaustudio_b200.camera import orbit_cameras
model, _, c = build_config("cfg3", K=1)
dev = torch.device("cuda")
cams = orbit_cameras(200, c["radius"], c["elev"], c["W"], c["H"], c["fovx"], c["fovy"], indices=list(range(60)))
hc = [bench.HostCamera(cm).upload(dev) for cm in cams]
model.to(dev).requires_grad_(True)
step = bench.make_step("new", model, dev, c["H"], c["W"], fused=1)
_C.set_pipelined(True)
for i in range(10): step(hc[i])
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(10, 60): step(hc[i])
th = time.perf_counter() - t; torch.cuda.synchronize(); tw = time.perf_counter() - t
print(f"host {th/50*1e3:.3f} ms/step, wall {tw/50*1e3:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for i in range(10, 60): step(hc[i])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
