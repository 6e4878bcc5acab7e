"""CPU oracle: structural properties, edge cases, independent cross-checks, f64 gradient check."""
import math

import numpy as np
import pytest

import scenes
from oracle import oracle as orc
from oracle.oracle import Oracle


def _fwd(s, dtype=np.float32):
    o = Oracle(dtype)
    out = o.forward(s["means3D"], s["opacities"], s["viewmatrix"], s["projmatrix"], s["campos"], s["tanfovx"],
                    s["tanfovy"], s["W"], s["H"], s["D"], shs=s.get("shs"), colors_precomp=s.get("colors_precomp"),
                    scales=s.get("scales"), rotations=s.get("rotations"), cov3D_precomp=s.get("cov3D_precomp"),
                    scale_modifier=s["scale_modifier"])
    return o, out


@pytest.mark.parametrize("case", "ABCD")
def test_binning_invariants(case):
    s = scenes.scene(case)
    o, out = _fwd(s)
    b, g = o.binning(), o.geometry()
    R = out["num_rendered"]
    assert R == int(g["tiles_touched"].sum()) == len(b["point_list"])
    rg = b["ranges"].astype(np.int64)
    n = rg[:, 1] - rg[:, 0]
    assert n.sum() == R and (n >= 0).all()
    nz = rg[n > 0]
    assert (nz[1:, 0] == nz[:-1, 1]).all() and nz[0, 0] == 0 and nz[-1, 1] == R  # ranges partition [0,R)
    depth_bits = g["depths"].view(np.uint32)
    for t in np.nonzero(n)[0][:200]:
        ids = b["point_list"][rg[t, 0]:rg[t, 1]].astype(np.int64)
        key = depth_bits[ids].astype(np.int64) * (1 << 32) + ids  # stable sort == total order (depth bits, index)
        assert (np.diff(key) > 0).all()
    # a Gaussian appears once per tile of its rect and nowhere else
    assert np.bincount(b["point_list"], minlength=len(g["tiles_touched"])).tolist() == g["tiles_touched"].tolist()
    assert ((out["radii"] > 0) == (g["tiles_touched"] > 0)).all()


@pytest.mark.parametrize("case", "ABCD")
def test_image_invariants(case):
    s = scenes.scene(case)
    o, out = _fwd(s)
    b = o.binning()
    np.testing.assert_allclose(out["opacity"][0], 1 - b["final_T"], atol=0, rtol=0)
    assert (out["opacity"] >= 0).all() and (out["opacity"] <= 1).all() and (b["final_T"] >= 1e-4 * 0.0).all()
    assert (out["color"] >= 0).all() and (out["depth"] >= 0).all()
    untouched = b["n_contrib"] == 0
    assert (out["median"][0][untouched] == 15.0).all()  # forward.cu:310
    assert (out["color"][:, untouched] == 0).all()       # no background blend (forward.cu:389-390)
    med_set = out["median"][1] > 0
    assert (out["median"][0][med_set] > 0.2).all() and (out["median"][2][med_set] == np.round(out["median"][2][med_set])).all()


def test_edge_cases():
    s = scenes.scene("A")
    # everything behind the camera -> nothing rendered, all radii 0
    s2 = dict(s); s2["means3D"] = s["means3D"] + 100 * (s["campos"] / np.linalg.norm(s["campos"]))
    o, out = _fwd(s2)
    assert out["num_rendered"] == 0 and (out["radii"] == 0).all() and (out["color"] == 0).all()
    assert (out["median"][0] == 15.0).all()
    g = o.backward(s["dL_color"], s["dL_depth"][0], s["dL_median"], s["dL_opacity"][0])
    assert all(np.abs(v).max() == 0 for v in g.values())
    # a single Gaussian, and P = 0
    s3 = {k: (v[:1] if isinstance(v, np.ndarray) and v.shape[:1] == (1500,) else v) for k, v in s.items()}
    o, out = _fwd(s3)
    assert out["num_rendered"] == int(o.geometry()["tiles_touched"].sum())
    s4 = {k: (v[:0] if isinstance(v, np.ndarray) and v.shape[:1] == (1500,) else v) for k, v in s.items()}
    o, out = _fwd(s4)
    assert out["num_rendered"] == 0 and out["color"].shape == (3, 64, 96)
    # mark_visible == near-plane test only (auxiliary.h:154)
    vis = orc.mark_visible(s["means3D"], s["viewmatrix"])
    pv = s["means3D"] @ s["viewmatrix"].reshape(4, 4)[:3, 2] + s["viewmatrix"].reshape(4, 4)[3, 2]
    assert (vis == (pv > 0.2)).mean() > 0.999


def test_sh_and_cov_against_independent_formulas():
    """SH colour vs the real-SH basis evaluated in float64 numpy; cov3D vs R S S^T R^T."""
    s = scenes.scene("A")
    o, out = _fwd(s)
    g = o.geometry()
    vis = out["radii"] > 0
    d = s["means3D"].astype(np.float64) - s["campos"].astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    sh = s["shs"].astype(np.float64)
    C = [0.28209479177387814, 0.4886025119029199,
         [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396],
         [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    basis = [C[0] + 0 * x, -C[1] * y, C[1] * z, -C[1] * x, C[2][0] * xy, C[2][1] * yz, C[2][2] * (2 * zz - xx - yy),
             C[2][3] * xz, C[2][4] * (xx - yy), C[3][0] * y * (3 * xx - yy), C[3][1] * xy * z,
             C[3][2] * y * (4 * zz - xx - yy), C[3][3] * z * (2 * zz - 3 * xx - 3 * yy), C[3][4] * x * (4 * zz - xx - yy),
             C[3][5] * z * (xx - yy), C[3][6] * x * (xx - 3 * yy)]
    rgb = np.maximum(sum(b * sh[:, k, :] for k, b in enumerate(basis)) + 0.5, 0)
    np.testing.assert_allclose(g["rgb"][vis], rgb[vis], atol=2e-5)
    R = scenes.quat_to_mat(s["rotations"].astype(np.float64))
    M = R * s["scales"].astype(np.float64)[:, None, :]
    Sg = M @ M.transpose(0, 2, 1)
    ref = np.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], 1)
    np.testing.assert_allclose(g["cov3D"][vis], ref[vis], rtol=1e-4, atol=1e-9)


def test_backward_is_the_gradient_of_forward_f64():
    """Central differences in float64 on the colour + depth loss (the opacity-map and median terms are
    deliberately non-gradient quirks 4/5 of the reference and are excluded)."""
    rng = np.random.default_rng(0)
    P, W, H = 50, 32, 32
    cam = scenes.camera(W, H, 40.0, (2.6, 0.0, 1.5))
    x = dict(means3D=rng.normal(0, 0.4, (P, 3)), opacities=rng.uniform(0.2, 0.9, (P, 1)),
             scales=np.exp(rng.normal(math.log(0.15), 0.3, (P, 3))), rotations=rng.normal(0, 1, (P, 4)),
             shs=rng.normal(0, 0.5, (P, 16, 3)))
    x["shs"][:, 0, :] += 1.0
    cst = dict(viewmatrix=cam.world_view_transform.numpy().astype(np.float64),
               projmatrix=cam.full_proj_transform.numpy().astype(np.float64),
               campos=cam.camera_center.numpy().astype(np.float64), tanfovx=math.tan(cam.FoVx * 0.5),
               tanfovy=math.tan(cam.FoVy * 0.5), W=W, H=H, sh_degree=3)
    wc, wd = rng.normal(0, 1, (3, H, W)), rng.normal(0, 1, (H, W))

    def loss(xx):
        o = Oracle(np.float64)
        out = o.forward(**xx, **cst)
        return (out["color"] * wc).sum() + (out["depth"][0] * wd).sum(), o
    _, o = loss(x)
    g = o.backward(wc, wd)
    for k in x:
        for _ in range(6):
            idx = tuple(rng.integers(0, n) for n in x[k].shape)
            xp = {a: b.copy() for a, b in x.items()}; xp[k][idx] += 1e-6
            xm = {a: b.copy() for a, b in x.items()}; xm[k][idx] -= 1e-6
            fd = (loss(xp)[0] - loss(xm)[0]) / 2e-6
            an = g[k][idx]
            assert abs(fd - an) <= 1e-5 * max(1.0, abs(fd), abs(an)), (k, idx, fd, an)


@pytest.mark.parametrize("bw,bh", [(8, 4), (16, 16)])
def test_cull_bound_is_conservative(bw, bh):
    """numpy restatement of may_touch() (gaustudio_b200/csrc/gsr_internal.cuh): the rectangle test -- 8x4 blocks in the
    compositing kernels, whole 16x16 tiles in the binning -- may only drop a Gaussian if NO pixel of the rectangle passes
    the reference's alpha >= 1/255 test."""
    rng = np.random.default_rng(3)
    n = 20000
    th = rng.uniform(0, np.pi, n); s1 = np.exp(rng.uniform(-1.0, 3.0, n)); s2 = np.exp(rng.uniform(-1.0, 3.0, n))
    c, s = np.cos(th), np.sin(th)
    cov = np.stack([c * c * s1 * s1 + s * s * s2 * s2, c * s * (s1 * s1 - s2 * s2), s * s * s1 * s1 + c * c * s2 * s2], 1)
    det = cov[:, 0] * cov[:, 2] - cov[:, 1] ** 2
    A, B, C = (cov[:, 2] / det).astype(np.float32), (-cov[:, 1] / det).astype(np.float32), (cov[:, 0] / det).astype(np.float32)
    o = rng.uniform(0.001, 1.0, n).astype(np.float32)
    gx, gy = rng.uniform(-40, 48, n).astype(np.float32), rng.uniform(-40, 44, n).astype(np.float32)
    rx0, ry0, rx1, ry1 = 0.0, 0.0, float(bw - 1), float(bh - 1)
    px, py = np.meshgrid(np.arange(bw, dtype=np.float32), np.arange(bh, dtype=np.float32))
    dx = gx[:, None, None] - px[None]; dy = gy[:, None, None] - py[None]
    power = -0.5 * (A[:, None, None] * dx * dx + C[:, None, None] * dy * dy) - B[:, None, None] * dx * dy
    alpha = np.minimum(0.99, o[:, None, None] * np.exp(power))
    touches = ((power <= 0) & (alpha >= 1 / 255)).any(axis=(1, 2))
    dxlo, dxhi, dylo, dyhi = gx - rx1, gx - rx0, gy - ry1, gy - ry0
    inx, iny = (dxlo <= 0) & (dxhi >= 0), (dylo <= 0) & (dyhi >= 0)
    dxe = np.where(dxlo > 0, dxlo, dxhi); dys = np.clip(-B * dxe / C, dylo, dyhi)
    q1 = A * dxe * dxe + 2 * B * dxe * dys + C * dys * dys
    dye = np.where(dylo > 0, dylo, dyhi); dxs = np.clip(-B * dye / A, dxlo, dxhi)
    q2 = A * dxs * dxs + 2 * B * dxs * dye + C * dye * dye
    qmin = np.minimum(np.where(inx, np.inf, q1), np.where(iny, np.inf, q2))
    mx, my = np.maximum(abs(dxlo), abs(dxhi)), np.maximum(abs(dylo), abs(dyhi))
    S = A * mx * mx + C * my * my + 2 * abs(B) * mx * my
    tau = 2 * np.log(255.0 * o)
    keep = (inx & iny) | ~(qmin > tau + 1e-5 * S + 1e-3)
    keep &= ~(o < 0.0039)
    assert not (touches & ~keep).any()          # never drops a contributing pair
    assert (keep & ~touches).mean() < 0.10      # and is reasonably tight
