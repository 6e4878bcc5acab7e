import math, sys
import torch
sys.path.insert(0, ".")
from gaustudio_b200.synthetic import build_config
from gaustudio_b200.rasterizer import GaussianRasterizationSettings
from gaustudio_b200 import _C
from oracle import ref_driver
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
model, cams, c = build_config(name, K=2)
dev = torch.device("cuda"); model.to(dev); cam = cams[0].to(dev)
xyz = model.get_attribute("xyz"); op = model.get_attribute("opacity"); sc = model.get_attribute("scale")
rot = model.get_attribute("rot"); sh = model.get_features.contiguous()
e = torch.Tensor([])
args = (torch.zeros(3, device=dev), xyz, e, op, sc, rot, 1.0, e, cam.world_view_transform, cam.full_proj_transform,
        math.tan(cam.FoVx*0.5), math.tan(cam.FoVy*0.5), cam.image_height, cam.image_width, sh, 3, cam.camera_center, False, False)
Rn, *_o, gN, bN, iN = _C.rasterize_gaussians(*args)
Rr, *_o2, gR, bR, iR = ref_driver.module().rasterize_gaussians(*args)
ref_radii = _o2[4]
print("R", Rn, Rr)
P = xyz.shape[0]
new = _C.debug_export(P, cam.image_width, cam.image_height, Rn, gN, bN, iN)
ref = ref_driver.parse_geometry(gR, P)
vis = ref_radii > 0
print("visible", int(vis.sum()), "radii equal", bool(torch.equal(ref_radii, _o[4])))
for k in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "tiles_touched", "clamped"):
    a, b = new[k][vis], ref[k][vis]
    nd = (a != b)
    print(k, "mismatch", int(nd.sum()), "/", nd.numel(), "maxabs", float((a.float()-b.float()).abs().max()))
pl = ref_driver.parse_binning(bR, Rr)
print("point_list equal:", bool(torch.equal(pl, new["point_list"])))
d = (new["rgb"][vis] != ref["rgb"][vis]).any(1)
print("rgb rows differing", int(d.sum()))
