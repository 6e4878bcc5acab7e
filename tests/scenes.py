"""Small seeded scenes shared by the golden generator and the parity tests (numpy only)."""
import math

import numpy as np

from gaustudio_b200.camera import look_at_camera


def camera(W, H, fov_deg=50.0, pos=(2.2, 1.4, 1.6)):
    fy = H / (2 * math.tan(math.radians(fov_deg) / 2))
    return look_at_camera(pos, (0, 0, 0), W, H, 2 * math.atan(W / (2 * fy)), math.radians(fov_deg))


def scene(case):
    """-> dict of float32 numpy inputs + settings.  Cases exercise the reference's option matrix."""
    rng = np.random.RandomState({"A": 11, "B": 12, "C": 13, "D": 14}[case])
    cfg = {"A": dict(P=1500, W=96, H=64, D=3, s0=0.05), "B": dict(P=800, W=80, H=80, D=1, s0=0.08),
           "C": dict(P=1200, W=70, H=50, D=0, s0=0.06), "D": dict(P=3000, W=160, H=112, D=2, s0=0.03)}[case]
    P = cfg["P"]
    d = rng.randn(P, 3); d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = d * (1.0 * rng.rand(P, 1) ** (1 / 3))
    scales = np.exp(math.log(cfg["s0"]) + 0.5 * rng.randn(P, 3))
    rot = rng.randn(P, 4); rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    opac = 1 / (1 + np.exp(-2.0 * rng.randn(P, 1)))
    shs = np.concatenate([rng.randn(P, 1, 3), 0.2 * rng.randn(P, 15, 3)], axis=1)
    s = dict(case=case, W=cfg["W"], H=cfg["H"], D=cfg["D"], scale_modifier=1.0, bg=np.zeros(3))
    if case == "C":
        # Gaussians behind / at the camera, one screen-filling splat, a degenerate flat one; modifier != 1; bg != 0
        xyz[:50] *= 6.0
        scales[50] = [1.5, 1.2, 0.9]; opac[50] = 0.6
        scales[51] = [1e-7, 0.2, 0.2]
        s["scale_modifier"] = 1.3
        s["bg"] = np.array([0.3, 0.1, 0.7])
    cam = camera(cfg["W"], cfg["H"])
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    s.update(means3D=f32(xyz), scales=f32(scales), rotations=f32(rot), opacities=f32(opac), shs=f32(shs),
             viewmatrix=cam.world_view_transform.numpy(), projmatrix=cam.full_proj_transform.numpy(),
             campos=cam.camera_center.numpy(), tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5))
    if case == "B":
        # precomputed colours and 3-D covariances (the other branch of both exactly-one-of rules), white bg
        R = quat_to_mat(rot)
        M = R * scales[:, None, :]
        Sg = M @ M.transpose(0, 2, 1)
        s["cov3D_precomp"] = f32(np.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], 1))
        s["colors_precomp"] = f32(rng.rand(P, 3))
        s["bg"] = np.ones(3)
        del s["shs"], s["scales"], s["rotations"]
    s["bg"] = f32(s["bg"])
    g = np.random.RandomState(99)
    s["dL_color"] = f32(g.randn(3, cfg["H"], cfg["W"])); s["dL_depth"] = f32(g.randn(1, cfg["H"], cfg["W"]))
    s["dL_median"] = f32(g.randn(3, cfg["H"], cfg["W"])); s["dL_opacity"] = f32(g.randn(1, cfg["H"], cfg["W"]))
    return s


def quat_to_mat(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


INPUT_KEYS = ("means3D", "scales", "rotations", "opacities", "shs", "colors_precomp", "cov3D_precomp")


def run_torch(s, rasterize, device, need_grad=True):
    """Run one forward(+backward) with a GaussianRasterizer-like callable `rasterize(rs, means3D, means2D,
    opacities, shs=, colors_precomp=, scales=, rotations=, cov3D_precomp=)`; returns numpy outputs / grads."""
    import torch
    from gaustudio_b200.rasterizer import GaussianRasterizationSettings
    t = {k: torch.tensor(s[k], device=device, requires_grad=need_grad) for k in INPUT_KEYS if k in s}
    m2d = torch.zeros_like(t["means3D"], requires_grad=need_grad)
    rs = GaussianRasterizationSettings(s["H"], s["W"], s["tanfovx"], s["tanfovy"], torch.tensor(s["bg"], device=device),
                                       s["scale_modifier"], torch.tensor(s["viewmatrix"], device=device),
                                       torch.tensor(s["projmatrix"], device=device), s["D"],
                                       torch.tensor(s["campos"], device=device), False, False)
    color, radii, depth, median, opacity = rasterize(
        rs, t["means3D"], m2d, t["opacities"], shs=t.get("shs"), colors_precomp=t.get("colors_precomp"),
        scales=t.get("scales"), rotations=t.get("rotations"), cov3D_precomp=t.get("cov3D_precomp"))
    out = dict(color=color, radii=radii, depth=depth, median=median, opacity=opacity)
    if need_grad:
        loss = sum((o * torch.tensor(s[k], device=device)).sum() for o, k in
                   ((color, "dL_color"), (depth, "dL_depth"), (median, "dL_median"), (opacity, "dL_opacity")))
        loss.backward()
        out["g_means2D"] = m2d.grad
        for k, v in t.items():
            out["g_" + k] = v.grad
    return {k: v.detach().cpu().numpy() for k, v in out.items() if v is not None}
