"""Small device ops beside the rasterizer: depth -> normal map (the vanilla path's "rendered normal")."""
import ctypes as C

import torch

from . import _lib


def depth2normal(depth, fx, fy, cx, cy, d_min=1e-3, d_max=100000.0, rot=None):
    """[H,W] (or [1,H,W]) float32 CUDA depth -> [H,W,3] normals; restates Camera.depth2normal with k=3
    (gaustudio/datasets/__init__.py:342-380) in one kernel.  Invalid pixels are (-1,-1,-1)."""
    if depth.dim() == 3:
        depth = depth[0]
    if not depth.is_cuda or depth.dtype != torch.float32:
        raise RuntimeError("depth2normal expects a float32 CUDA tensor")
    depth = depth.contiguous()
    H, W = depth.shape
    out = torch.empty(H, W, 3, dtype=torch.float32, device=depth.device)
    rp = None
    if rot is not None:
        rot = rot.to(device=depth.device, dtype=torch.float32).contiguous()
        rp = C.c_void_p(rot.data_ptr())
    with torch.cuda.device(depth.device):
        rc = _lib.lib().gsr_depth2normal(C.c_void_p(depth.data_ptr()), W, H, float(fx), float(fy), float(cx),
                                         float(cy), float(d_min), float(d_max), rp, C.c_void_p(out.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream(depth.device).cuda_stream))
    if rc < 0:
        raise RuntimeError("gsr_depth2normal failed: " + _lib.last_error())
    return out


def depth2point(depth, fx, fy, cx, cy, cam_to_world=None):
    """[H,W] float32 CUDA depth -> [H,W,3] points in camera coordinates, or world coordinates when
    `cam_to_world` (4x4, inverse of the extrinsics) is given; restates Camera.depth2point
    (gaustudio/datasets/__init__.py:307-339)."""
    if depth.dim() == 3:
        depth = depth[0]
    if not depth.is_cuda or depth.dtype != torch.float32:
        raise RuntimeError("depth2point expects a float32 CUDA tensor")
    depth = depth.contiguous()
    H, W = depth.shape
    out = torch.empty(H, W, 3, dtype=torch.float32, device=depth.device)
    mp = None
    if cam_to_world is not None:
        cam_to_world = cam_to_world.to(device=depth.device, dtype=torch.float32).contiguous()
        mp = C.c_void_p(cam_to_world.data_ptr())
    with torch.cuda.device(depth.device):
        rc = _lib.lib().gsr_depth2point(C.c_void_p(depth.data_ptr()), W, H, float(fx), float(fy), float(cx), float(cy),
                                        mp, C.c_void_p(out.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream(depth.device).cuda_stream))
    if rc < 0:
        raise RuntimeError("gsr_depth2point failed: " + _lib.last_error())
    return out
