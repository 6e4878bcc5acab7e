"""BASELINE cfg 5 (5M Gaussians, 1440x1080, forward-only depth/median/opacity/normal pass): correctness
invariants at full size + per-view time (exact and pipelined forward), at D=3 and at D=0 with M=16 (quirk 13)."""
import sys, math, time, ctypes, torch
sys.path.insert(0, ".")
from gaustudio_b200 import _C, _lib, renderers
from gaustudio_b200.synthetic import build_config
t = time.time(); model, cams, c = build_config("cfg5", K=12); print("scene built", round(time.time() - t, 1), "s")
dev = torch.device("cuda"); model.to(dev)
P, W, H = c["P"], c["W"], c["H"]
L = _lib.lib()
for D in (3, 0):
    model.active_sh_degree = D
    for fused in (False, True):
        r = renderers.make({"name": "vanilla_renderer", "fused_activations": fused})
        _C.set_pipelined(True)
        with torch.no_grad():
            for i in range(3):
                cam = cams[i].to(dev); out = r.render(cam, model); n = cam.depth2normal(out["rendered_depth"][0])
            torch.cuda.synchronize(); L.gsr_profile_enable(1)
            t = time.time()
            for i in range(3, 11):
                cam = cams[i].to(dev); out = r.render(cam, model); n = cam.depth2normal(out["rendered_depth"][0])
            torch.cuda.synchronize(); dt = (time.time() - t) / 8
        _C.check_pipeline(wait=True)
        ms = (ctypes.c_float * 8)(); cn = (ctypes.c_int * 8)(); L.gsr_profile_read(ms, cn); L.gsr_profile_enable(0)
        names = ["pre", "scan", "scatter", "sort", "render", "rbwd", "pbwd", "normal"]
        print(f"D={D} fused={fused}: {dt*1e3:.2f} ms/view ({1/dt:.1f} views/s)", {n_: round(ms[i] / max(cn[i], 1), 3) for i, n_ in enumerate(names) if cn[i]},
              "mem GB", round(torch.cuda.max_memory_allocated() / 2**30, 2))
_C.set_pipelined(False)
# invariants on one view (exact mode)
model.active_sh_degree = 3
cam = cams[0].to(dev); e = torch.Tensor([])
with torch.no_grad():
    R, color, depth, median, opac, radii, gb, bb, ib = _C.rasterize_gaussians(
        torch.zeros(3, device=dev), model.get_attribute("xyz"), e, model.get_attribute("opacity"), model.get_attribute("scale"),
        model.get_attribute("rot"), 1.0, e, cam.world_view_transform, cam.full_proj_transform, math.tan(cam.FoVx * .5),
        math.tan(cam.FoVy * .5), H, W, model.get_features.contiguous(), 3, cam.camera_center, False, False)
ex = _C.debug_export(P, W, H, R, gb, bb, ib)
rg = ex["ranges"].long(); n = rg[:, 1] - rg[:, 0]
ids = ex["point_list"].long(); key = (ex["depths"].view(torch.int32).long()[ids] << 32) | ids
tile_of = torch.repeat_interleave(torch.arange(rg.shape[0], device=dev), n)
ok = bool(((key[1:] > key[:-1]) | (tile_of[1:] != tile_of[:-1])).all())
print("R", R, "visible", int((radii > 0).sum()), "max tile n", int(n.max()), "tiles > 8192:", int((n > 8192).sum()), "sorted:", ok,
      "sum ok:", int(n.sum()) == R == int(ex["tiles_touched"].long().sum()), "opacity==1-T:", bool(torch.equal(opac[0], 1 - ex["final_T"])))
