"""First-contact GPU check: new kernels vs the compiled reference vs the CPU oracle on a small scene."""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gaustudio_b200.synthetic import build_config
from gaustudio_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
from gaustudio_b200 import _C
from oracle import ref_driver
from oracle.oracle import Oracle

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
kw = {}
if len(sys.argv) > 2: kw["P"] = int(sys.argv[2])
model, cams, c = build_config(name, K=2, **kw)
dev = torch.device("cuda")
model.to(dev); cam = cams[0].to(dev)
def inputs(requires_grad):
    xyz = model.get_attribute("xyz").detach().clone().requires_grad_(requires_grad)
    op = model.get_attribute("opacity").detach().clone().requires_grad_(requires_grad)
    sc = model.get_attribute("scale").detach().clone().requires_grad_(requires_grad)
    rot = model.get_attribute("rot").detach().clone().requires_grad_(requires_grad)
    sh = model.get_features.detach().clone().requires_grad_(requires_grad)
    m2d = torch.zeros_like(xyz, requires_grad=requires_grad)
    return xyz, m2d, op, sh, sc, rot
rs = GaussianRasterizationSettings(cam.image_height, cam.image_width, math.tan(cam.FoVx*0.5), math.tan(cam.FoVy*0.5),
        torch.zeros(3, device=dev), 1.0, cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center, False, True)
H, W = cam.image_height, cam.image_width
g = torch.Generator(device="cpu").manual_seed(7)
wc = torch.randn(3, H, W, generator=g).to(dev); wd = torch.randn(1, H, W, generator=g).to(dev)
wo = torch.randn(1, H, W, generator=g).to(dev); wm = torch.randn(3, H, W, generator=g).to(dev)

def run(kind):
    xyz, m2d, op, sh, sc, rot = inputs(True)
    if kind == "new":
        color, radii, depth, median, opac = GaussianRasterizer(rs)(xyz, m2d, op, shs=sh, scales=sc, rotations=rot)
    else:
        color, radii, depth, median, opac = ref_driver.rasterize(rs, xyz, m2d, op, shs=sh, scales=sc, rotations=rot)
    loss = (color*wc).sum() + (depth*wd).sum() + (opac*wo).sum() + (median*wm).sum()
    loss.backward()
    torch.cuda.synchronize()
    return dict(color=color, radii=radii, depth=depth, median=median, opac=opac,
                g_xyz=xyz.grad, g_m2d=m2d.grad, g_op=op.grad, g_sh=sh.grad, g_sc=sc.grad, g_rot=rot.grad)
new = run("new"); print("new ok")
ref = run("ref"); print("ref ok")
def cmp(a, b, name):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    d = (a-b).abs(); den = b.abs().max().item() + 1e-20
    print(f"{name:8s} maxabs {d.max().item():.3e}  rel-to-max {d.max().item()/den:.3e}  nonzero-diff {int((d>0).sum())}/{d.numel()}  bit-equal {bool(torch.equal(a,b))}")
for k in new: cmp(new[k], ref[k], k)
# oracle
o = Oracle()
xyz, m2d, op, sh, sc, rot = [t.detach().cpu().numpy() for t in inputs(False)]
out = o.forward(xyz, op, cam.world_view_transform.cpu().numpy(), cam.full_proj_transform.cpu().numpy(), cam.camera_center.cpu().numpy(),
                rs.tanfovx, rs.tanfovy, W, H, 3, shs=sh, scales=sc, rotations=rot)
print("oracle R", out["num_rendered"])
for k, kk in (("color","color"),("depth","depth"),("median","median"),("opacity","opac")):
    cmp(torch.from_numpy(out[k]), ref[kk], "orc-"+k)
print("radii equal oracle/ref:", int((torch.from_numpy(out["radii"]) != ref["radii"].cpu()).sum()))
gb = o.backward(wc.cpu().numpy(), wd.cpu().numpy()[0], wm.cpu().numpy(), wo.cpu().numpy()[0])
for k, kk in (("means3D","g_xyz"),("means2D","g_m2d"),("opacities","g_op"),("shs","g_sh"),("scales","g_sc"),("rotations","g_rot")):
    cmp(torch.from_numpy(gb[k]), ref[kk], "orc-"+kk)
