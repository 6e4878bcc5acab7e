"""Runs a few cfg3 steps (plugin render + loss + backward + depth2normal); used under ncu."""
import sys, math, torch
sys.path.insert(0, ".")
from gaustudio_b200 import renderers, ops
from gaustudio_b200.synthetic import build_config
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, cams, c = build_config(name, K=8)
dev = torch.device("cuda"); model.to(dev).requires_grad_(True)
r = renderers.make({"name": "vanilla_renderer", "fused_activations": True})
for i in range(steps):
    cam = cams[i].to(dev)
    out = r.render(cam, model)
    loss = out["render"].abs().mean() + 0.1 * out["rendered_depth"].abs().mean() + 0.1 * out["rendered_final_opacity"].abs().mean()
    loss.backward()
    K = cam.intrinsics
    n = ops.depth2normal(out["rendered_depth"].detach()[0], float(K[0,0]), float(K[1,1]), float(K[0,2]), float(K[1,2]))
torch.cuda.synchronize()
print("done", float(loss))
