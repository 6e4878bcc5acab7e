"""Host-side time per step (no sync) vs GPU time, for combinations of pipelined / profile / fused."""
import sys, time, ctypes, torch
sys.path.insert(0, ".")
import bench
from gaustudio_b200 import _C, _lib
from gaustudio_b200.synthetic import build_config
from gaustudio_b200.camera import orbit_cameras
model, _, c = build_config("cfg3", K=1)
dev = torch.device("cuda")
cams = orbit_cameras(200, c["radius"], c["elev"], c["W"], c["H"], c["fovx"], c["fovy"], indices=list(range(40)))
hc = [bench.HostCamera(cm).upload(dev) for cm in cams]
model.to(dev).requires_grad_(True)
L = _lib.lib()
for fused in (0, 1):
    for pipe in (0, 1):
        for prof in (0, 1):
            step = bench.make_step("new", model, dev, c["H"], c["W"], fused=fused)
            _C.set_pipelined(bool(pipe)); L.gsr_profile_enable(prof)
            for i in range(5): step(hc[i])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t = time.perf_counter(); e0.record()
            for i in range(5, 35): step(hc[i])
            e1.record(); th = time.perf_counter() - t
            torch.cuda.synchronize(); tw = time.perf_counter() - t
            ms = (ctypes.c_float * 8)(); cn = (ctypes.c_int * 8)(); L.gsr_profile_read(ms, cn); L.gsr_profile_enable(0)
            if pipe: _C.check_pipeline(wait=True)
            print(f"fused {fused} pipelined {pipe} profile {prof}: host {th/30*1e3:.3f} ms/step  wall {tw/30*1e3:.3f}  gpu-events {e0.elapsed_time(e1)/30:.3f}")
# split of host time inside one step (fused=1, pipelined)
import math
from gaustudio_b200 import renderers, ops
import torch.nn.functional as F
r = renderers.make({"name": "vanilla_renderer", "fused_activations": True}); _C.set_pipelined(True)
tc = torch.rand(3, c["H"], c["W"], device=dev); td = torch.rand(1, c["H"], c["W"], device=dev)
acc = [0.0]*5
for i in range(5, 35):
    cam = hc[i]
    t0 = time.perf_counter(); out = r.render(cam, model)
    t1 = time.perf_counter(); loss = F.l1_loss(out["render"], tc) + 0.1 * F.l1_loss(out["rendered_depth"], td) + 0.1 * F.l1_loss(out["rendered_final_opacity"], td)
    t2 = time.perf_counter(); loss.backward()
    t3 = time.perf_counter(); n = ops.depth2normal(out["rendered_depth"].detach()[0], cam.fx, cam.fy, cam.cx, cam.cy)
    t4 = time.perf_counter()
    for p in model.parameters_list(): p.grad = None
    t5 = time.perf_counter()
    for k, v in enumerate((t1-t0, t2-t1, t3-t2, t4-t3, t5-t4)): acc[k] += v
torch.cuda.synchronize()
print("host split ms: render %.3f loss %.3f backward %.3f normal %.3f zero %.3f" % tuple(a/30*1e3 for a in acc))
