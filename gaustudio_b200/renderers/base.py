"""BaseRenderer.render: the L3 plugin call of the reference (`gaustudio/renderers/base.py:10-63`).

Gathers the activated attributes, builds GaussianRasterizationSettings, calls GaussianRasterizer and returns
the same 9-key dict.  Differences: no `plyfile` import (absent here, and irrelevant to rendering), and the
screen-space tensor is created on the model's device instead of the literal "cuda".
"""
import math

import torch

from ..rasterizer import GaussianRasterizationSettings, GaussianRasterizer


class BaseRenderer:
    def render(self, viewpoint_camera, gaussian_model):
        xyz, shs, colors_precomp, opacity, scales, rotations, cov3D_precomp = \
            self.get_gaussians_properties(viewpoint_camera, gaussian_model)
        # zero tensor through which autograd returns the 2-D (screen-space) mean gradients
        screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
        tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
        tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
        raster_settings = GaussianRasterizationSettings(
            image_height=int(viewpoint_camera.image_height),
            image_width=int(viewpoint_camera.image_width),
            tanfovx=tanfovx,
            tanfovy=tanfovy,
            bg=self.bg_color,
            scale_modifier=self.scaling_modifier,
            viewmatrix=viewpoint_camera.world_view_transform,
            projmatrix=viewpoint_camera.full_proj_transform,
            sh_degree=gaussian_model.active_sh_degree if shs is not None else 1,
            campos=viewpoint_camera.camera_center,
            prefiltered=False,
            debug=self.debug)
        rasterizer = GaussianRasterizer(raster_settings=raster_settings)
        rendered_image, radii, rendered_depth, rendered_median_map, rendered_final_opacity = rasterizer(
            means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp, opacities=opacity,
            scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp)
        return {"render": rendered_image,
                "rendered_depth": rendered_depth,
                "rendered_median_depth": rendered_median_map[0:1],
                "rendered_median_weight": rendered_median_map[1:2],
                "rendered_median_id": rendered_median_map[2:3].int(),
                "viewspace_points": screenspace_points,
                "visibility_filter": radii > 0,
                "rendered_final_opacity": rendered_final_opacity,
                "radii": radii}
