"""Does the un-fused forward ever read its torch-computed inputs (activations, cat(f_dc, f_rest)) stale?
Two different models of the same shape are rendered alternately, every render with freshly computed activation tensors
(so their blocks hold the OTHER model's values from the previous iteration), and compared bit-exactly with a baseline
rendered with a device synchronisation between the torch ops and the forward.
phases: aligned SH tensor (per-thread TMA rows) / 4-byte-misaligned SH tensor (plain loads).
usage: stale_hunt.py [iters]      (GSR_PRIORITY=0/1 is read once per process by the library)"""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from gaustudio_b200 import _C
from gaustudio_b200.synthetic import make_unbounded_scene, build_config

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
P, W, H = 40000, 240, 180
dev = torch.device("cuda")
_, cams, _ = build_config("cfg5", P=1000, K=1, W=W, H=H)
cam = cams[0].to(dev)
models = [make_unbounded_scene(P, seed).to(dev) for seed in (5, 6)]
e = torch.Tensor([]).to(dev)
bg = torch.zeros(3, device=dev)
vm, pm, cc = cam.world_view_transform.contiguous(), cam.full_proj_transform.contiguous(), cam.camera_center.contiguous()
tfx, tfy = math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5)


def forward(m, misalign, sync):
    op = torch.sigmoid(m._opacity)
    sc = torch.exp(m._scale)
    rot = F.normalize(m._rot)
    if misalign:
        buf = torch.empty(P * 48 + 8, device=dev)
        sh = buf[1:1 + P * 48].view(P, 16, 3)
        sh[:, :1] = m._f_dc
        sh[:, 1:] = m._f_rest
    else:
        sh = torch.cat((m._f_dc, m._f_rest), dim=1)
    if sync:
        torch.cuda.synchronize()
    out = _C.rasterize_gaussians(bg, m._xyz, e, op, sc, rot, 1.0, e, vm, pm, tfx, tfy, H, W, sh, 3, cc, False, False)
    return out[1].clone(), out[5].clone(), out[2].clone()   # colour, radii, depth


with torch.no_grad():
    _C.set_speculation(False)
    base = {}
    for mi, m in enumerate(models):
        for mis in (False, True):
            r = [forward(m, mis, True) for _ in range(3)]
            torch.cuda.synchronize()
            if not all(torch.equal(r[0][k], r[j][k]) for j in (1, 2) for k in range(3)):
                print("baseline does not reproduce", mi, mis, flush=True)
            base[(mi, mis)] = r[0]
    print("aligned vs misaligned SH baselines equal:", torch.equal(base[(0, False)][0], base[(0, True)][0]), flush=True)
    print(f"GSR_PRIORITY={os.environ.get('GSR_PRIORITY', '1')} baselines ok", flush=True)
    for spec in (False, True):
        _C.set_speculation(spec)
        for mis in (False, True):
            for idle_start in (False, True):
                bad_col = bad_geo = 0
                first = None
                for it in range(iters):
                    for mi, m in enumerate(models):
                        if idle_start:
                            torch.cuda.synchronize()
                        col, rad, dep = forward(m, mis, False)
                        b = base[(mi, mis)]
                        if not torch.equal(rad, b[1]) or not torch.equal(dep, b[2]):
                            bad_geo += 1
                            first = first or (it, mi, "geometry", int((rad != b[1]).sum()), float((dep - b[2]).abs().max()))
                        elif not torch.equal(col, b[0]):
                            bad_col += 1
                            first = first or (it, mi, "colour", int((col != b[0]).sum()), float((col - b[0]).abs().max()))
                print(f"  speculation={spec} sh={'misaligned/plain loads' if mis else 'aligned/TMA rows'} idle_start={idle_start}: "
                      f"{bad_geo} geometry + {bad_col} colour-only mismatches of {2 * iters}; first {first}", flush=True)
