"""A few cfg3 steps of the UNMODIFIED reference extension (oracle/_ref/_refC.so): its torch activations -> renderCUDA
forward -> loss -> backward.  Run under ncu (-k regex:renderCUDA) to put the reference's own K4 / K5 counters on
record next to ours.  Usage: python tools/profile_ref.py [cfg3] [steps]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from gaustudio_b200.synthetic import build_config  # noqa: E402
from oracle import ref_driver, ref_torch_ops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model, cams, c = build_config(name, K=8)
dev = torch.device("cuda")
model.to(dev).requires_grad_(True)
bg = torch.zeros(3, device=dev)
for i in range(steps):
    cam = cams[i].to(dev)
    xyz, shs, opacity, scales, rotations = ref_torch_ops.gaussian_properties(model)
    rs = ref_driver.RefSettings(c["H"], c["W"], math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), bg, 1.0,
                                cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, False)
    color, radii, depth, median, opac = ref_driver.rasterize(rs, xyz, torch.zeros_like(xyz, requires_grad=True) + 0, opacity,
                                                             shs=shs, scales=scales, rotations=rotations)
    loss = color.abs().mean() + 0.1 * depth.abs().mean() + 0.1 * opac.abs().mean()
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
