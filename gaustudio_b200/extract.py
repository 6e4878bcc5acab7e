"""Extraction post-pass on the GPU -- the step right after the rasterizer in gaustudio/scripts/extract_pcd.py
(SURVEY.md 8f row 2).  Same function names, argument meaning and results as the reference script:

  masked_bilateral_filter(depth_map, mask, d, sigma_color, sigma_space)      extract_pcd.py:185-238
  normal_fusion(pcd, all_ids_list, all_normals_list, all_confidences_list, cameras)   extract_pcd.py:108-183
  getNerfppNorm(cameras)                                                     gaustudio/datasets/utils.py:82-104
  extract_view(camera, render_pkg, scene_radius)                             the loop body, extract_pcd.py:314-337
  extract_pcd(renderer, pcd, cameras)                                        the loop + fusion, extract_pcd.py:309-345

The reference moves every depth map to the host for OpenCV (dilate + bilateralFilter) and back; here the filter,
the normal extraction and the fusion passes are CUDA kernels behind the C ABI (include/gsr.h) and nothing is
synchronised per view.  The k-nearest-neighbour search of the final smoothing (scipy's cKDTree on the host in the
reference) is a uniform-grid search on the device as well (`knn`, gsr_knn_grid).
"""
import ctypes as C

import math

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{what} must be a CUDA tensor (the extraction post-pass has no CPU path)")


def _extrinsics_host(camera):
    """float32 numpy copy of camera.extrinsics without touching the device when the camera keeps R / T on the host
    (world_view_transform = [[R^T, T], [0, 1]] up to the recentring of datasets/__init__.py:52-64, which only moves
    the translation)."""
    wvt = getattr(camera, "world_view_transform", None)
    if isinstance(wvt, torch.Tensor) and not wvt.is_cuda:
        return wvt.transpose(0, 1).contiguous().numpy().astype(np.float32)
    if getattr(camera, "trans", None) is not None and np.allclose(camera.trans, 0) and getattr(camera, "scale", 1.0) == 1.0:
        E = np.eye(4, dtype=np.float32)
        E[:3, :3] = np.asarray(camera.R, np.float32).T
        E[:3, 3] = np.asarray(camera.T, np.float32)
        return E
    return camera.extrinsics.detach().cpu().numpy().astype(np.float32)


def masked_bilateral_filter(depth_map, mask, d=3, sigma_color=75, sigma_space=75):
    """Bilateral-filter the valid part of a depth map; returns (filtered_depth, new_mask) where new_mask marks
    pixels whose d x d window is entirely valid.  [H,W] float32 / bool CUDA tensors in and out."""
    _need_cuda(depth_map, "depth_map")
    if depth_map.dim() != 2:
        raise RuntimeError("masked_bilateral_filter expects an [H,W] depth map")
    depth = depth_map.detach().float().contiguous()
    m = mask.detach().to(device=depth.device, dtype=torch.bool).contiguous()
    H, W = depth.shape
    out = torch.empty_like(depth)
    new_mask = torch.empty(H, W, dtype=torch.bool, device=depth.device)
    scratch = torch.empty(2, dtype=torch.int32, device=depth.device)
    with torch.cuda.device(depth.device):
        rc = _lib.lib().gsr_masked_bilateral(_ptr(depth), _ptr(m), W, H, int(d), float(sigma_color), float(sigma_space),
                                             _ptr(out), _ptr(new_mask), _ptr(scratch), _stream(depth.device))
    if rc < 0:
        raise RuntimeError("gsr_masked_bilateral failed: " + _lib.last_error())
    return out, new_mask.to(mask.dtype)


def getNerfppNorm(cameras):
    """Scene centre / radius from the camera centres (gaustudio/datasets/utils.py:82-104)."""
    centres = []
    for cam in cameras:
        w2c = np.eye(4)
        w2c[:3, :3] = np.asarray(cam.R, np.float64).T
        w2c[:3, 3] = np.asarray(cam.T, np.float64)
        centres.append(np.linalg.inv(w2c)[:3, 3:4])
    centres = np.hstack(centres)
    centre = centres.mean(axis=1, keepdims=True)
    dist = np.linalg.norm(centres - centre, axis=0)
    return {"translate": -centre.flatten(), "radius": float(dist.max() * 1.1), "min_radius": float(dist.min() * 1.5)}


def _intrinsics_host(camera):
    """(fx, fy, cx, cy) as host floats without touching the device: from the camera's FoV / size / principal point
    (the formula of Camera.intrinsics, gaustudio/datasets/__init__.py:229-237) when it has them."""
    if all(hasattr(camera, a) for a in ("FoVx", "FoVy", "image_width", "image_height")):
        W, H = int(camera.image_width), int(camera.image_height)
        pp = getattr(camera, "principal_point_ndc", None)
        px, py = (0.5, 0.5) if pp is None else (float(pp[0]), float(pp[1]))
        return (W / (2.0 * math.tan(camera.FoVx / 2.0)), H / (2.0 * math.tan(camera.FoVy / 2.0)), W * px, H * py)
    K = camera.intrinsics  # a device-resident K costs one synchronising read here
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


def extract_view(camera, render_pkg, scene_radius, d=3, sigma_color=75, sigma_space=75, compact=True):
    """One view of the extraction loop.  Returns dict(filtered_depth, fg_mask, cam_normals [H,W,3], valid [H,W],
    ids [n] int64, normals [n,3] (negated world normals), confidences [n]).

    compact=True gathers the valid pixels like the reference (`ids[valid]`: a boolean-mask gather, i.e. one stream
    synchronisation per view).  compact=False is sync-free: ids / normals / confidences cover ALL H*W pixels and the
    invalid ones carry id -1, which `normal_fusion` skips (negative ids mean "no observation" there)."""
    opacity = render_pkg["rendered_final_opacity"][0].contiguous()
    depth = render_pkg["rendered_depth"][0]
    median_depth = render_pkg["rendered_median_depth"][0].contiguous()
    median_ids = render_pkg["rendered_median_id"][0]
    _need_cuda(opacity, "render_pkg tensors")
    dev = opacity.device
    H, W = opacity.shape
    filtered, fg = masked_bilateral_filter(depth, opacity > 0.1, d, sigma_color, sigma_space)
    fx, fy, cx, cy = _intrinsics_host(camera)
    rot = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(_extrinsics_host(camera)[:3, :3]).T)).to(dev)
    cam_normals = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    neg_world = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(H, W, dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().gsr_extract_normals(_ptr(filtered), _ptr(fg), _ptr(opacity), _ptr(median_depth), W, H,
                                            float(fx), float(fy), float(cx), float(cy), _ptr(rot),
                                            float(scene_radius * 0.8), 0.5, _ptr(cam_normals), _ptr(neg_world),
                                            _ptr(valid), _stream(dev))
    if rc < 0:
        raise RuntimeError("gsr_extract_normals failed: " + _lib.last_error())
    res = {"filtered_depth": filtered, "fg_mask": fg, "cam_normals": cam_normals, "valid": valid}
    if compact:
        res.update(ids=median_ids[valid].long(), normals=neg_world[valid], confidences=opacity[valid])
    else:
        res.update(ids=torch.where(valid, median_ids.long(), -1).reshape(-1), normals=neg_world.reshape(-1, 3),
                   confidences=opacity.reshape(-1))
    return res


def _fusion_pass(xyz, ids_list, normals_list, conf_list, cam_ts, mean, sums, weights, touched):
    L = _lib.lib()
    dev = xyz.device
    with torch.cuda.device(dev):
        for ids, n, c, t in zip(ids_list, normals_list, conf_list, cam_ts):
            if ids.numel() == 0:
                continue
            ids = ids.to(device=dev, dtype=torch.int64).contiguous()
            n = n.to(device=dev, dtype=torch.float32).contiguous()
            c = c.to(device=dev, dtype=torch.float32).contiguous()
            rc = L.gsr_normal_fusion_pass(ids.numel(), _ptr(ids), _ptr(n), _ptr(c), xyz.shape[0], _ptr(xyz),
                                          t[0], t[1], t[2], _ptr(mean), 0.8, _ptr(sums), _ptr(weights), _ptr(touched),
                                          _stream(dev))
            if rc < 0:
                raise RuntimeError("gsr_normal_fusion_pass failed: " + _lib.last_error())
        out = torch.empty_like(sums)
        if L.gsr_normal_fusion_mean(xyz.shape[0], _ptr(sums), _ptr(weights), _ptr(out), _stream(dev)) < 0:
            raise RuntimeError("gsr_normal_fusion_mean failed: " + _lib.last_error())
    return out


def normal_fusion(pcd, all_ids_list, all_normals_list, all_confidences_list, cameras, smooth=True):
    """Weighted, consistency-checked fusion of per-view normals into one normal per observed Gaussian.
    Returns (unique_ids ascending, normals [n,3]) like the reference.
    Ids outside [0, P) are skipped: negative ids are the "no observation" sentinel of `extract_view(compact=False)`.
    (The reference indexes `pcd._xyz[ids]` with them: -1 wraps to the last Gaussian and ids >= P raise; its own loop
    never produces either, extract_pcd.py:325-337.)"""
    xyz = pcd._xyz.detach().float().contiguous()
    _need_cuda(xyz, "pcd._xyz")
    dev, P = xyz.device, xyz.shape[0]
    # what the reference uses as the camera position: extrinsics[:3, 3] (taken from the host copy of R / T)
    cam_ts = [_extrinsics_host(cam)[:3, 3].tolist() for cam in cameras]
    sums = torch.zeros(P, 3, device=dev); weights = torch.zeros(P, device=dev)
    touched = torch.zeros(P, dtype=torch.uint8, device=dev)
    mean = _fusion_pass(xyz, all_ids_list, all_normals_list, all_confidences_list, cam_ts, None, sums, weights, touched)
    sums.zero_(); weights.zero_()
    mean = _fusion_pass(xyz, all_ids_list, all_normals_list, all_confidences_list, cam_ts, mean, sums, weights, None)
    unique_ids = torch.nonzero(touched, as_tuple=False).flatten()
    mean = mean[unique_ids]
    if not smooth or unique_ids.numel() == 0:
        return unique_ids, mean
    # spatial smoothing over the 10 nearest surface points (extract_pcd.py:170-181: scipy cKDTree on the host there);
    # here the neighbour search stays on the device (the reference needs >= 10 surface points)
    dist, idx = knn(xyz[unique_ids], k=min(10, int(unique_ids.numel())))
    w = torch.exp(-dist.double() / 0.1)
    sm = (mean[idx].double() * w.unsqueeze(-1)).sum(1).float()
    return unique_ids, torch.nn.functional.normalize(sm, p=2, dim=1)


_KNN_K = (1, 4, 8, 10, 16)


def knn(points, k=10):
    """k nearest neighbours of every point among `points` [n,3] (float32, CUDA) -> (dist [n,k] float32, idx [n,k] int64),
    ascending distance with the point itself first -- what `cKDTree(points).query(points, k)` returns
    (extract_pcd.py:170-171), without leaving the device.  Uniform-grid search (gsr_knn_grid): the points are sorted by
    cell with torch ops, the bounding box / cell size never come back to the host."""
    _need_cuda(points, "points")
    pts = points.detach().float().contiguous()
    n, dev = pts.shape[0], pts.device
    if k not in _KNN_K or n <= k:  # tiny inputs / unusual k: exact all-pairs search (still on the device)
        d = torch.cdist(pts.double(), pts.double())
        dist, idx = d.topk(min(k, n), dim=1, largest=False)
        return dist.float(), idx
    G = min(256, max(8, 1 << math.ceil(math.log2(max(n, 2) ** (1.0 / 3.0)))))
    lo, hi = pts.min(dim=0).values, pts.max(dim=0).values
    ext = (hi - lo).clamp_min(1e-12)
    inv_h = 1.0 / (ext.max() / G * 1.0001 + 1e-12)
    dims = torch.clamp(torch.ceil(ext * inv_h), min=1.0, max=float(G))
    cell = torch.minimum(torch.clamp(((pts - lo) * inv_h).floor(), min=0.0), dims - 1.0)
    cid = ((cell[:, 2] * dims[1] + cell[:, 1]) * dims[0] + cell[:, 0]).long()
    order = torch.argsort(cid, stable=True)
    sorted_pts = pts[order].contiguous()
    counts = torch.zeros(G ** 3 + 1, dtype=torch.int64, device=dev)
    counts.scatter_add_(0, cid + 1, torch.ones_like(cid))
    cell_start = torch.cumsum(counts, 0).to(torch.int32).contiguous()
    grid = torch.cat([lo, inv_h.reshape(1), dims, torch.zeros(1, device=dev)]).float().contiguous()
    out_i = torch.empty(n, k, dtype=torch.int32, device=dev)
    out_d = torch.empty(n, k, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().gsr_knn_grid(n, k, _ptr(sorted_pts), _ptr(cell_start), _ptr(grid), _ptr(out_i), _ptr(out_d),
                                     _stream(dev))
    if rc < 0:
        raise RuntimeError("gsr_knn_grid failed: " + _lib.last_error())
    dist = torch.empty_like(out_d)
    idx = torch.empty(n, k, dtype=torch.int64, device=dev)
    dist[order] = out_d                     # rows back to the caller's order ...
    idx[order] = order[out_i.long()]        # ... and neighbour indices back to the caller's numbering
    return dist, idx


def SH2RGB(sh):
    """gaustudio/utils/sh_utils.py:114-118."""
    return sh * 0.28209479177387814 + 0.5


def extract_pcd(renderer, pcd, cameras, d=3, sigma_color=75, sigma_space=75, smooth=True):
    """Render every camera and fuse a surface point cloud: (xyz [n,3], rgb [n,3], normals [n,3], per-view results).
    The loop of extract_pcd.py:309-345 without its image / .cam file output and without meshing."""
    scene_radius = getNerfppNorm(cameras)["radius"]
    ids_l, nrm_l, conf_l, views = [], [], [], []
    for camera in cameras:
        with torch.no_grad():
            pkg = renderer.render(camera, pcd)
        v = extract_view(camera, pkg, scene_radius, d, sigma_color, sigma_space, compact=False)  # no per-view sync
        ids_l.append(v["ids"]); nrm_l.append(v["normals"]); conf_l.append(v["confidences"])
        views.append(v)
    unique_ids, normals = normal_fusion(pcd, ids_l, nrm_l, conf_l, cameras, smooth=smooth)
    f_dc = pcd._f_dc[unique_ids]
    rgb = SH2RGB(f_dc.reshape(f_dc.shape[0], -1)[:, :3]).clip(0, 1)
    return pcd._xyz[unique_ids], rgb, normals, views
