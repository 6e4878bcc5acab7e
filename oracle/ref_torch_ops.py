"""Torch restatement of the reference's Python-side ops that sit beside the rasterizer in BASELINE cfg 3/5:
Camera.depth2point / depth2normal (gaustudio/datasets/__init__.py:106-112, 307-380).  TEST INFRASTRUCTURE
(reference arm of bench.py and the parity tests); our own code, written from the behavioural description."""
import torch
import torch.nn.functional as F


def depth2point_camera(depth, K):
    H, W = depth.shape
    x = torch.arange(W, dtype=torch.float32, device=depth.device) / (W - 1)
    y = torch.arange(H, dtype=torch.float32, device=depth.device) / (H - 1)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    scale = torch.tensor([[W - 1, H - 1]], device=depth.device)
    xy = torch.stack([xx, yy], dim=-1) * scale * depth[..., None]
    xyz = torch.cat([xy, depth[..., None]], dim=-1)
    return xyz @ torch.inverse(K.to(depth.device).t())


def depth2normal(depth, K, d_min=1e-3, d_max=100000.0, rot=None):
    H, W = depth.shape
    pts = depth2point_camera(depth, K).permute(2, 0, 1)[None]
    pad = F.pad(pts, (1, 1, 1, 1), mode="constant", value=0)
    valid = ((pad[:, 2:] > d_min) & (pad[:, 2:] < d_max)).float()
    vert = pad[:, :, :H, 1:1 + W] - pad[:, :, 2:2 + H, 1:1 + W]
    hori = pad[:, :, 1:1 + H, :W] - pad[:, :, 1:1 + H, 2:2 + W]
    vm = (valid[:, :, 1:1 + H, 1:1 + W] * valid[:, :, :H, 1:1 + W] * valid[:, :, 2:2 + H, 1:1 + W] *
          valid[:, :, 1:1 + H, :W] * valid[:, :, 1:1 + H, 2:2 + W]) > 0.5
    n = F.normalize(-torch.linalg.cross(vert, hori, dim=1), p=2.0, dim=1, eps=1e-12)
    if rot is not None:
        n = (n.permute(0, 2, 3, 1) @ rot.to(n.device)).permute(0, 3, 1, 2)
    n[~vm.repeat(1, 3, 1, 1)] = -1
    return n.squeeze(0).permute(1, 2, 0)


def gaussian_properties(model):
    """Per-view attribute activations of the reference's vanilla renderer
    (gaustudio/renderers/vanilla_renderer.py:28-52 with both *_python options off): exp / sigmoid / normalize through
    the model's `get_attribute`, SH as cat(f_dc, f_rest).  -> (xyz, shs, opacity, scales, rotations)"""
    return (model.get_attribute("xyz"), model.get_features, model.get_attribute("opacity"),
            model.get_attribute("scale"), model.get_attribute("rot"))
