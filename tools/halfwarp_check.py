"""Round-2 A/B of the experimental half-warp compositing kernels (GSR_HALFWARP=1, gsr_render.cu template HALF):
runs the same views in two subprocesses (the flag is read once per process), compares outputs and gradients, and
prints per-kernel times.  Expected: forward outputs BIT-IDENTICAL (per-pixel order is unchanged, only culled pairs
differ), gradients equal up to float-atomic ordering (<= 1e-3 relative).  Usage: python tools/halfwarp_check.py [cfg3]"""
import json, os, subprocess, sys, tempfile

CHILD = r'''
import sys, ctypes, torch
sys.path.insert(0, ".")
from gaustudio_b200 import renderers, _lib
from gaustudio_b200.synthetic import build_config
name, out = sys.argv[1], sys.argv[2]
model, cams, c = build_config(name, K=6); dev = torch.device("cuda"); model.to(dev).requires_grad_(True)
r = renderers.make({"name": "vanilla_renderer", "fused_activations": True}); L = _lib.lib()
res = {}
for i in range(6):
    if i == 3: torch.cuda.synchronize(); L.gsr_profile_enable(1)
    o = r.render(cams[i].to(dev), model)
    loss = o["render"].abs().mean() + 0.1 * o["rendered_depth"].abs().mean() + 0.1 * o["rendered_final_opacity"].abs().mean()
    loss.backward()
    if i == 0:
        res = {k: o[k].detach().cpu() for k in ("render", "rendered_depth", "rendered_median_depth", "rendered_final_opacity", "radii")}
        res["grads"] = [p.grad.detach().cpu().clone() for p in model.parameters_list()]
torch.cuda.synchronize()
ms = (ctypes.c_float * 8)(); cn = (ctypes.c_int * 8)(); L.gsr_profile_read(ms, cn)
names = ["preprocess_fwd", "tile_scan", "scatter", "tile_sort", "render_fwd", "render_bwd", "preprocess_bwd", "depth2normal"]
res["ms"] = {n: ms[i] / max(cn[i], 1) for i, n in enumerate(names) if cn[i]}
torch.save(res, out)
'''

import torch
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
outs = {}
with tempfile.TemporaryDirectory() as d:
    for flag in ("0", "1"):
        path = os.path.join(d, f"r{flag}.pt")
        env = dict(os.environ, GSR_HALFWARP=flag)
        subprocess.run([sys.executable, "-c", CHILD, name, path], check=True, env=env, timeout=600)
        outs[flag] = torch.load(path)
a, b = outs["0"], outs["1"]
rep = {"config": name, "forward_bit_identical": all(torch.equal(a[k], b[k]) for k in a if k not in ("grads", "ms"))}
rel = []
for ga, gb in zip(a["grads"], b["grads"]):
    rel.append(float((ga - gb).abs().max() / ga.abs().max().clamp_min(1e-30)))
rep["grad_max_rel_diff"] = max(rel)
rep["ms_default"] = {k: round(v, 4) for k, v in a["ms"].items()}
rep["ms_halfwarp"] = {k: round(v, 4) for k, v in b["ms"].items()}
rep["ok"] = rep["forward_bit_identical"] and rep["grad_max_rel_diff"] < 1e-3
print(json.dumps(rep))
