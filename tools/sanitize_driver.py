"""Forward + backward of a small scene straight through the binding (gaustudio_b200._C: no autograd, no loss kernels),
un-fused and fused, so that everything compute-sanitizer reports comes from this library's kernels.
Usage: compute-sanitizer --tool <tool> python tools/sanitize_driver.py"""
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from gaustudio_b200 import _C  # noqa: E402
from gaustudio_b200.synthetic import build_config  # noqa: E402

# SAN_SCENE="cfg2,150000,320,320": another named config / size (denser tiles: every sort tier and long compositing lists)
_scene = os.environ.get("SAN_SCENE", "cfg1").split(",")
_kw = dict(P=int(_scene[1]), W=int(_scene[2]), H=int(_scene[3])) if len(_scene) == 4 else {}
model, cams, c = build_config(_scene[0], K=2, **_kw)
dev = torch.device("cuda")
model.to(dev)
H, W, P = c["H"], c["W"], c["P"]
e = torch.Tensor([])
g = torch.Generator().manual_seed(0)
dL = [torch.randn(s, H, W, generator=g).to(dev) for s in (3, 1, 3, 1)]
bg = torch.zeros(3, device=dev)
with torch.no_grad():
    for cam in cams:
        cam.to(dev)
        tail = (cam.world_view_transform, cam.full_proj_transform, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5))
        xyz, op, sc, rot, sh = (model.get_attribute("xyz"), model.get_attribute("opacity"), model.get_attribute("scale"),
                                model.get_attribute("rot"), model.get_features.contiguous())
        R, color, depth, median, opac, radii, gb, bb, ib = _C.rasterize_gaussians(
            bg, xyz, e, op, sc, rot, 1.0, e, *tail, H, W, sh, 3, cam.camera_center, False, False)
        grads = _C.rasterize_gaussians_backward(bg, xyz, radii, e, sc, rot, 1.0, e, *tail, *dL, sh, 3, cam.camera_center, gb, R,
                                                bb, ib, False)
        f_dc, f_rest = model._f_dc.reshape(P, -1, 3).contiguous(), model._f_rest.reshape(P, -1, 3).contiguous()
        out = _C.rasterize_gaussians(bg, model._xyz, e, model._opacity, model._scale, model._rot, 1.0, e, *tail, H, W, e, 3,
                                     cam.camera_center, False, False, _fused=(f_dc, f_rest))
        fg = _C.rasterize_gaussians_fused_backward(bg, model._xyz, out[5], f_dc, f_rest, model._opacity, model._scale,
                                                   model._rot, 1.0, *tail, *dL, 3, cam.camera_center, out[6], out[0], out[7],
                                                   out[8], False)
        ex = _C.debug_export(P, W, H, R, gb, bb, ib)
torch.cuda.synchronize()
print("done", R, float(color.sum()), float(sum(t.double().abs().sum() for t in grads)), float(sum(t.double().abs().sum() for t in fg)))
