"""Hunt for run-to-run differences of the (deterministic) forward: the same view rendered over and over, fused and
un-fused, speculation on/off, with allocator / timing perturbation in between.  Any difference between two renders of the
same inputs is a race.  usage: race_hunt.py [P] [iters]   (GSR_SCAN_CLUSTER / GSR_SPECULATE are read by the library)"""
import os, sys, math, random
import torch
sys.path.insert(0, ".")
from gaustudio_b200 import _C, renderers
from gaustudio_b200.synthetic import build_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
W, H = 240, 180
dev = torch.device("cuda")
model, cams, _ = build_config("cfg5", P=P, K=1, W=W, H=H)
model.to(dev)
cam = cams[0].to(dev)
rr = {f: renderers.make({"name": "vanilla_renderer", "fused_activations": f}) for f in (False, True)}
KEYS = ("render", "rendered_depth", "rendered_median_depth", "rendered_final_opacity")
e = torch.Tensor([]).to(dev)


def tiles_of(mask2d):
    t = mask2d.view(-1, H, W).any(0)
    Hp, Wp = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    pad = torch.zeros(Hp, Wp, dtype=torch.bool, device=dev); pad[:H, :W] = t
    return pad.view(Hp // 16, 16, Wp // 16, 16).any(3).any(1).flatten().nonzero().flatten().tolist()


def tile_sizes():
    out = _C.rasterize_gaussians(torch.zeros(3, device=dev), model.get_attribute("xyz"), e, model.get_attribute("opacity"),
                                 model.get_attribute("scale"), model.get_attribute("rot"), 1.0, e, cam.world_view_transform,
                                 cam.full_proj_transform, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), H, W,
                                 model.get_features.contiguous(), 3, cam.camera_center, False, False)
    ex = _C.debug_export(P, W, H, out[0], out[6], out[7], out[8])
    return (ex["ranges"][:, 1] - ex["ranges"][:, 0]).tolist(), ex["point_list"].clone(), ex["ranges"].clone()


with torch.no_grad():
    sizes, pl0, rg0 = tile_sizes()
    print(f"P={P} cluster={os.environ.get('GSR_SCAN_CLUSTER','1')} tiles: max {max(sizes)} over2048 {sum(s > 2048 for s in sizes)} "
          f"over6144 {sum(s > 6144 for s in sizes)} over12288 {sum(s > 12288 for s in sizes)}", flush=True)
    base = {}
    bad = 0
    junk = []
    random.seed(0)
    for it in range(iters):
        spec = it % 3 != 0
        _C.set_speculation(spec)
        for f in (False, True):
            out = rr[f].render(cam, model)
            cur = [out[k].clone() for k in KEYS]
            if f not in base:
                base[f] = cur
                continue
            diff = torch.zeros(H, W, dtype=torch.bool, device=dev)
            for a, b in zip(cur, base[f]):
                diff |= (a != b).view(-1, H, W).any(0)
            if diff.any():
                bad += 1
                tl = tiles_of(diff)
                print(f"  MISMATCH it={it} fused={f} spec={spec}: {int(diff.sum())} px in tiles {tl[:12]} sizes {[sizes[t] for t in tl[:12]]}", flush=True)
        # the sorted list itself
        s2, pl, rg = tile_sizes()
        if not (torch.equal(pl, pl0) and torch.equal(rg, rg0)):
            bad += 1
            if pl.numel() != pl0.numel():
                print(f"  LIST LENGTH MISMATCH it={it} spec={spec}: {pl.numel()} vs {pl0.numel()}", flush=True)
            else:
                d = (pl != pl0).nonzero().flatten()
                owner = torch.repeat_interleave(torch.arange(len(sizes), device=dev), torch.tensor(sizes, device=dev))
                tl = sorted(set(owner[d].tolist()))
                print(f"  LIST MISMATCH it={it} spec={spec}: {d.numel()} entries, tiles {tl[:12]} sizes {[sizes[t] for t in tl[:12]]} "
                      f"ranges equal {torch.equal(rg, rg0)}", flush=True)
        # perturb allocator state and timing
        junk.append(torch.empty(random.randint(1, 64) << 18, device=dev))
        if len(junk) > 6:
            del junk[random.randrange(len(junk))]
        if it % 4 == 1:
            torch.cuda.synchronize()
        if it % 5 == 2:
            (torch.randn(1 << random.randint(10, 24), device=dev) * 2).sum()
    print(f"done: {bad} mismatching renders of {iters * 3}; speculation stats {_C.speculation_stats()}", flush=True)
