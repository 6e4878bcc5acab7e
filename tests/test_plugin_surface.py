"""Drop-in surface: module names, Settings fields, exceptions and registry behaviour of the reference
($RAST/gaustudio_diff_gaussian_rasterization/__init__.py:160-223, gaustudio/renderers/__init__.py:1-28)."""
import pytest
import torch


def test_module_names_and_exports():
    import diff_gaussian_rasterization as d
    import gaustudio_diff_gaussian_rasterization as g
    for m in (g, d):
        assert m.GaussianRasterizer.__name__ == "GaussianRasterizer"
        assert hasattr(m, "rasterize_gaussians") and hasattr(m, "_C")
        for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
            assert hasattr(m._C, fn)
    assert g.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def _rast():
    from gaustudio_diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    z = torch.zeros(4, 4)
    return GaussianRasterizer(GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, z, z, 0, torch.zeros(3),
                                                            False, False))


def test_exactly_one_of_rules():
    r = _rast()
    P = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(P, P, torch.zeros(4, 1), scales=P, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(P, P, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=P, scales=P, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(P, P, torch.zeros(4, 1), colors_precomp=P)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(P, P, torch.zeros(4, 1), colors_precomp=P, scales=P, cov3D_precomp=torch.zeros(4, 6))


def test_bad_means_shape_and_cpu_tensors_fail_loudly():
    r = _rast()
    with pytest.raises(RuntimeError, match="num_points, 3"):
        r(torch.zeros(4, 2), torch.zeros(4, 2), torch.zeros(4, 1), colors_precomp=torch.zeros(4, 3),
          cov3D_precomp=torch.zeros(4, 6))
    # no CPU fallback: CPU inputs are an error, not a slow path
    with pytest.raises(RuntimeError, match="CUDA"):
        r(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 1), colors_precomp=torch.zeros(4, 3),
          cov3D_precomp=torch.zeros(4, 6))


def test_registry():
    from gaustudio_b200 import renderers
    with pytest.raises(ValueError, match="required"):
        renderers.make({})
    with pytest.raises(ValueError, match="Unknown renderer"):
        renderers.make("nope")
    r1 = renderers.make("vanilla_renderer")
    r2 = renderers.make({"name": "vanilla_renderer", "white_background": True, "scaling_modifier": 0.5})
    assert r1.scaling_modifier == 1.0 and r1.bg_color.tolist() == [0, 0, 0] and not r1.bg_color.is_cuda
    assert r2.scaling_modifier == 0.5 and r2.bg_color.tolist() == [1, 1, 1]

    @renderers.register("dummy")
    class Dummy:
        def __init__(self, cfg):
            self.cfg = cfg
    assert isinstance(renderers.make({"name": "dummy", "x": 1}), Dummy)


def test_vanilla_yaml_keys_accepted():
    # gaustudio/configs/vanilla.yaml:22-28
    from gaustudio_b200 import renderers
    r = renderers.make({"name": "vanilla_renderer", "scaling_modifier": 1., "white_background": False,
                        "convert_SHs_python": False, "compute_cov3D_python": False, "debug": False})
    assert r.debug is False


def test_extract_surface_fails_loudly_without_cuda():
    """The post-pass mirrors extract_pcd.py's function names / arguments and has no CPU path."""
    import inspect
    import math
    import types

    import pytest
    import torch

    from gaustudio_b200 import extract
    from gaustudio_b200.camera import orbit_cameras
    assert list(inspect.signature(extract.masked_bilateral_filter).parameters) == \
        ["depth_map", "mask", "d", "sigma_color", "sigma_space"]                      # extract_pcd.py:185
    assert list(inspect.signature(extract.normal_fusion).parameters)[:5] == \
        ["pcd", "all_ids_list", "all_normals_list", "all_confidences_list", "cameras"]  # extract_pcd.py:108
    sig = inspect.signature(extract.masked_bilateral_filter).parameters
    assert (sig["d"].default, sig["sigma_color"].default, sig["sigma_space"].default) == (3, 75, 75)
    with pytest.raises(RuntimeError, match="no CPU path"):
        extract.masked_bilateral_filter(torch.rand(8, 8), torch.ones(8, 8, dtype=torch.bool))
    with pytest.raises(RuntimeError, match="no CPU path"):
        extract.normal_fusion(types.SimpleNamespace(_xyz=torch.rand(4, 3)), [], [], [], [])
    cams = orbit_cameras(6, 3.0, 30.0, 64, 48, 0.8, 0.6)
    norm = extract.getNerfppNorm(cams)   # datasets/utils.py:82-104: radius = 1.1 * max distance to the mean centre
    assert abs(norm["radius"] - 1.1 * 3.0 * math.cos(math.radians(30.0))) < 1e-6 and norm["translate"].shape == (3,)
