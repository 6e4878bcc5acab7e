"""Driver for the UNMODIFIED reference CUDA extension built by oracle/build_ref.py (oracle/_ref/_refC.so).

TEST INFRASTRUCTURE ONLY (tests/, bench.py --impl reference, __graft_entry__).  The .so is the reference's own
code compiled where it lies; this file is our own thin autograd wrapper around its three pybind functions
(`rasterize_gaussians`, `rasterize_gaussians_backward`, `mark_visible`; $RAST/ext.cpp:15-19) -- the reference's
Python package is not copied.  Needs a GPU to run.
"""
import importlib.util
import os
from typing import NamedTuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "_refC.so")
_mod = None


def available():
    return os.path.exists(SO)


def module():
    global _mod
    if _mod is None:
        if not available():
            raise ImportError("oracle/_ref/_refC.so missing: run `python oracle/build_ref.py` where /root/reference exists")
        spec = importlib.util.spec_from_file_location("_refC", SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


class RefSettings(NamedTuple):
    """Field-for-field the reference's GaussianRasterizationSettings ($RAST/.../__init__.py:160-172), so the reference
    arm of bench.py needs nothing from gaustudio_b200's rasterizer module."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class RefRasterize(torch.autograd.Function):
    """Argument packing of the reference's _RasterizeGaussians ($RAST/.../__init__.py:44-158)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        C = module()
        (num_rendered, color, depth, median, opacity, radii, geom, binning, img) = C.rasterize_gaussians(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree,
            rs.campos, rs.prefiltered, rs.debug)
        ctx.rs, ctx.num_rendered = rs, num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, median, opacity

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_median, g_opacity):
        C = module()
        rs = ctx.rs
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img = ctx.saved_tensors
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot) = C.rasterize_gaussians_backward(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, g_color, g_depth, g_median, g_opacity, sh, rs.sh_degree, rs.campos,
            geom, ctx.num_rendered, binning, img, rs.debug)
        return g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rot, g_cov3D, None


def rasterize(rs, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
              cov3D_precomp=None):
    """Same call as GaussianRasterizer.forward; `rs.bg` must be a CUDA tensor if backward is used
    (the reference dereferences it on the device, backward.cu:586)."""
    e = torch.Tensor([])
    return RefRasterize.apply(means3D, means2D, e if shs is None else shs,
                              e if colors_precomp is None else colors_precomp, opacities,
                              e if scales is None else scales, e if rotations is None else rotations,
                              e if cov3D_precomp is None else cov3D_precomp, rs)


def parse_geometry(geom, P):
    """Views into the reference's geomBuffer: GeometryState::fromChunk, rasterizer_impl.cu:155-171
    (each array aligned to 128 B from the chunk's address, rasterizer_impl.h:23-29)."""
    base = geom.data_ptr()
    off = 0

    def take(nbytes):
        nonlocal off
        start = ((base + off + 127) & ~127) - base
        off = start + nbytes
        return geom[start:start + nbytes]
    out = {}
    out["depths"] = take(4 * P).view(torch.float32)
    out["clamped"] = take(3 * P).view(torch.uint8).reshape(P, 3)
    out["radii"] = take(4 * P).view(torch.int32)
    out["means2D"] = take(8 * P).view(torch.float32).reshape(P, 2)
    out["cov3D"] = take(24 * P).view(torch.float32).reshape(P, 6)
    out["conic_opacity"] = take(16 * P).view(torch.float32).reshape(P, 4)
    out["rgb"] = take(12 * P).view(torch.float32).reshape(P, 3)
    out["tiles_touched"] = take(4 * P).view(torch.int32)
    return out


def parse_binning(binning, R):
    """point_list of BinningState::fromChunk (rasterizer_impl.cu:182-194): first array of the chunk."""
    base = binning.data_ptr()
    start = ((base + 127) & ~127) - base
    return binning[start:start + 4 * R].view(torch.int32)


def parse_image_ranges(img, num_pixels, num_tiles):
    """Per-tile [start, end) of ImageState::fromChunk (rasterizer_impl.cu:173-180): accum_alpha f32[N], n_contrib
    u32[N], ranges uint2[N] (allocated for N = W*H entries, the first `num_tiles` are used), each 128-byte aligned."""
    base = img.data_ptr()
    off = 0

    def take(nbytes):
        nonlocal off
        start = ((base + off + 127) & ~127) - base
        off = start + nbytes
        return img[start:start + nbytes]
    take(4 * num_pixels)
    take(4 * num_pixels)
    return take(8 * num_pixels).view(torch.int32).reshape(num_pixels, 2)[:num_tiles]
