"""BaseRenderer.render: the L3 plugin call of the reference (`gaustudio/renderers/base.py:10-63`).

Contract kept: subclasses provide `get_gaussians_properties(camera, model)` ->
(xyz, shs, colors_precomp, opacity, scales, rotations, cov3D_precomp) plus the attributes `bg_color`,
`scaling_modifier`, `debug`; `render` returns the reference's 9-key dict.  Differences: no `plyfile` import
(absent here, irrelevant to rendering) and the screen-space tensor lives on the model's device instead of the
literal "cuda".
"""
import math

import torch

from ..rasterizer import GaussianRasterizationSettings, GaussianRasterizer

RESULT_KEYS = ("render", "rendered_depth", "rendered_median_depth", "rendered_median_weight", "rendered_median_id",
               "viewspace_points", "visibility_filter", "rendered_final_opacity", "radii")


def settings_for(camera, *, bg, scale_modifier, sh_degree, debug):
    """GaussianRasterizationSettings of one view (field meaning: $RAST/.../__init__.py:160-172)."""
    return GaussianRasterizationSettings(
        image_height=int(camera.image_height), image_width=int(camera.image_width),
        tanfovx=math.tan(camera.FoVx * 0.5), tanfovy=math.tan(camera.FoVy * 0.5), bg=bg,
        scale_modifier=scale_modifier, viewmatrix=camera.world_view_transform,
        projmatrix=camera.full_proj_transform, sh_degree=sh_degree, campos=camera.camera_center,
        prefiltered=False, debug=debug)


def pack_result(image, radii, depth, median_map, opacity, screenspace):
    """The median map's three channels are (depth, weight, Gaussian id as float): forward.cu:392-394."""
    values = (image, depth, median_map[0:1], median_map[1:2], median_map[2:3].int(), screenspace, radii > 0, opacity,
              radii)
    return dict(zip(RESULT_KEYS, values))


class BaseRenderer:
    def render(self, viewpoint_camera, gaussian_model):
        xyz, shs, colors_precomp, opacity, scales, rotations, cov3D_precomp = \
            self.get_gaussians_properties(viewpoint_camera, gaussian_model)
        # a zero tensor in the graph: autograd hands back the 2-D (screen-space) mean gradients through it
        screenspace = torch.zeros_like(xyz, requires_grad=True) + 0
        try:
            screenspace.retain_grad()
        except Exception:
            pass
        degree = gaussian_model.active_sh_degree if shs is not None else 1
        rs = settings_for(viewpoint_camera, bg=self.bg_color, scale_modifier=self.scaling_modifier, sh_degree=degree,
                          debug=self.debug)
        image, radii, depth, median_map, final_opacity = GaussianRasterizer(raster_settings=rs)(
            means3D=xyz, means2D=screenspace, opacities=opacity, shs=shs, colors_precomp=colors_precomp,
            scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
        return pack_result(image, radii, depth, median_map, final_opacity, screenspace)
