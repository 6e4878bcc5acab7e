import numpy as np
import torch


def new_rasterize(rs, means3D, means2D, opacities, **kw):
    from gaustudio_b200.rasterizer import GaussianRasterizer
    return GaussianRasterizer(rs)(means3D, means2D, opacities, **kw)


def ref_rasterize(rs, means3D, means2D, opacities, **kw):
    from oracle import ref_driver
    return ref_driver.rasterize(rs, means3D, means2D, opacities, **kw)


def oracle_run(s):
    """CPU oracle outputs + grads in the naming of scenes.run_torch."""
    from oracle.oracle import Oracle
    o = Oracle()
    out = o.forward(s["means3D"], s["opacities"], s["viewmatrix"], s["projmatrix"], s["campos"], s["tanfovx"],
                    s["tanfovy"], s["W"], s["H"], s["D"], shs=s.get("shs"), colors_precomp=s.get("colors_precomp"),
                    scales=s.get("scales"), rotations=s.get("rotations"), cov3D_precomp=s.get("cov3D_precomp"),
                    scale_modifier=s["scale_modifier"])
    g = o.backward(s["dL_color"], s["dL_depth"][0], s["dL_median"], s["dL_opacity"][0], bg=s["bg"])
    r = dict(color=out["color"], radii=out["radii"], depth=out["depth"], median=out["median"], opacity=out["opacity"],
             num_rendered=out["num_rendered"], g_means2D=g["means2D"], g_means3D=g["means3D"],
             g_opacities=g["opacities"])
    if "shs" in s:
        r["g_shs"] = g["shs"]
    else:
        r["g_colors_precomp"] = g["colors_precomp"]
    if "scales" in s:
        r["g_scales"], r["g_rotations"] = g["scales"], g["rotations"]
    else:
        r["g_cov3D_precomp"] = g["cov3D_precomp"]
    return r


def assert_grads_close(a, b, rel=1e-3, floor=1e-4, what=""):
    """BASELINE: <= 1e-3 relative on gradients, with an absolute floor (relative to the tensor's scale)
    because the reference's float atomics make its own gradients order-dependent."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = np.abs(b).max()
    err = np.abs(a - b)
    bound = rel * np.abs(b) + floor * scale + 1e-30
    bad = err > bound
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.size} beyond tolerance, worst {err.max():.3e} (scale {scale:.3e})"


def assert_images_close(a, b, atol=1e-4, outlier_frac=0.0, what=""):
    err = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    frac = (err > atol).mean()
    assert frac <= outlier_frac, f"{what}: {frac:.2e} of pixels beyond {atol} (max {err.max():.3e})"
