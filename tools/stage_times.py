"""Per-kernel ms for a few cfg3 views (forward only unless bwd=1)."""
import sys, ctypes, math, torch
sys.path.insert(0, ".")
from gaustudio_b200 import _C, _lib, renderers
from gaustudio_b200.synthetic import build_config
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
model, cams, c = build_config(name, K=8)
dev = torch.device("cuda"); model.to(dev)
r = renderers.make({"name": "vanilla_renderer", "fused_activations": True})
L = _lib.lib()
with torch.no_grad():
    for i in range(3): r.render(cams[i].to(dev), model)
    torch.cuda.synchronize(); L.gsr_profile_enable(1)
    for i in range(3, 8): r.render(cams[i].to(dev), model)
    torch.cuda.synchronize()
ms = (ctypes.c_float * 8)(); cn = (ctypes.c_int * 8)(); L.gsr_profile_read(ms, cn)
names = ["preprocess_fwd", "tile_scan", "scatter", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd", "depth2normal"]
print({n: round(ms[i] / max(cn[i], 1), 4) for i, n in enumerate(names) if cn[i]})
