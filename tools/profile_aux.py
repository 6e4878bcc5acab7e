"""One launch each (after one warm-up) of the kernels added after the main ncu capture, at cfg-3 scale; used under ncu:
un-fused projection fwd/bwd (per-row TMA staging of the [P,16,3] SH tensor), the extraction post-pass, fused AdamW."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
from gaustudio_b200 import renderers, extract, _lib
from gaustudio_b200.optimizers import FusedAdam
from gaustudio_b200.synthetic import build_config
dev = torch.device("cuda")
model, cams, c = build_config("cfg3", K=4); model.to(dev).requires_grad_(True)
r = renderers.make({"name": "vanilla_renderer"})  # reference op sequence: activations in torch, [P,16,3] SH tensor
for i in range(2):
    out = r.render(cams[i].to(dev), model)
    (out["render"].abs().mean() + 0.1 * out["rendered_depth"].abs().mean()).backward()
radius = extract.getNerfppNorm(cams)["radius"]
with torch.no_grad():
    for i in range(2):
        v = extract.extract_view(cams[i].to(dev), {k: (t.detach() if isinstance(t, torch.Tensor) else t) for k, t in out.items()}, radius)
P, n = 1_000_000, 1_000_000
pc = type("P", (), {})(); pc._xyz = torch.randn(P, 3, device=dev)
ids = torch.randint(0, P, (n,), device=dev); nrm = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=1); conf = torch.rand(n, device=dev)
cam = type("Cam", (), {})(); cam.extrinsics = torch.eye(4)
extract.normal_fusion(pc, [ids], [nrm], [conf], [cam], smooth=False)
big = torch.nn.Parameter(torch.randn(59_000_000, device=dev)); big.grad = torch.randn_like(big)
fa = FusedAdam([big], lr=1e-3, eps=1e-15)
fa.step(); fa.step(zero_grad=True)
torch.cuda.synchronize(); print("done")
