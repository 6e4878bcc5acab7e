"""Generates tests/golden/extract_golden.npz by RUNNING the reference's own post-pass functions in this container.

`gaustudio/scripts/extract_pcd.py` cannot be imported (open3d / trimesh / omegaconf are absent), so the two
function definitions the fixtures need -- `masked_bilateral_filter` (:185-238) and `normal_fusion` (:108-183) --
are pulled out of the unmodified reference source with `ast` at generation time and executed against the real
cv2 (4.13), torch and scipy.  Nothing of the reference is written into the repo; only inputs and outputs are.

    python tests/golden/make_golden_extract.py        (needs /root/reference, cv2, scipy; CPU only)
"""
import ast
import os
import types

import cv2
import numpy as np
import torch
from scipy.spatial import cKDTree

REF = "/root/reference/gaustudio/scripts/extract_pcd.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "extract_golden.npz")


def reference_functions():
    tree = ast.parse(open(REF).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("masked_bilateral_filter", "normal_fusion")]
    assert len(keep) == 2
    ns = {"np": np, "cv2": cv2, "torch": torch, "cKDTree": cKDTree}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns["masked_bilateral_filter"], ns["normal_fusion"]


def depth_case(seed, H, W, holes):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    depth = 2.5 + 0.6 * xx - 0.3 * yy * yy + 0.05 * torch.randn(H, W, generator=g)
    depth = depth + (xx > 0.3).float() * 1.2  # a depth discontinuity
    opacity = torch.rand(H, W, generator=g) * 0.2 + 0.8
    for _ in range(holes):
        cy, cx = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, W, (1,), generator=g))
        r = int(torch.randint(1, 5, (1,), generator=g))
        opacity[max(cy - r, 0):cy + r, max(cx - r, 0):cx + r] = 0.02
    opacity[:, :2] = 0.05  # an invalid band touching the image border
    return depth.float().contiguous(), opacity.float().contiguous()


def main():
    bilateral, fusion = reference_functions()
    out = {}
    for name, (seed, H, W, holes, d) in {"a": (11, 48, 64, 6, 3), "b": (12, 40, 56, 3, 5), "c": (13, 33, 47, 0, 3)}.items():
        depth, opacity = depth_case(seed, H, W, holes)
        mask = opacity > 0.1
        f, m = bilateral(depth, mask, d=d)
        out[f"bil_{name}_depth"] = depth.numpy(); out[f"bil_{name}_mask"] = mask.numpy()
        out[f"bil_{name}_d"] = np.int32(d)
        out[f"bil_{name}_filtered"] = f.numpy(); out[f"bil_{name}_newmask"] = m.numpy()
    # non-default sigmas, where the range kernel actually bites
    depth, opacity = depth_case(14, 48, 64, 4)
    mask = opacity > 0.1
    f, m = bilateral(depth, mask, d=3, sigma_color=0.05, sigma_space=1.5)
    out["bil_s_depth"] = depth.numpy(); out["bil_s_mask"] = mask.numpy(); out["bil_s_d"] = np.int32(3)
    out["bil_s_sigma"] = np.array([0.05, 1.5], np.float32)
    out["bil_s_filtered"] = f.numpy(); out["bil_s_newmask"] = m.numpy()

    # normal fusion: P Gaussians, V views; several pixels per Gaussian and per view
    g = torch.Generator().manual_seed(21)
    P, V = 400, 4
    xyz = torch.randn(P, 3, generator=g)
    true_n = torch.nn.functional.normalize(xyz + 0.2 * torch.randn(P, 3, generator=g), dim=1)
    cams, ids_l, nrm_l, conf_l = [], [], [], []
    for v in range(V):
        E = torch.eye(4)
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        E[:3, :3] = q
        E[:3, 3] = torch.randn(3, generator=g) * 3.0
        cams.append(types.SimpleNamespace(extrinsics=E))
        n = int(torch.randint(500, 900, (1,), generator=g))
        ids = torch.randint(0, P - 40, (n,), generator=g)  # the last 40 Gaussians are never seen
        nrm = torch.nn.functional.normalize(true_n[ids] + 0.3 * torch.randn(n, 3, generator=g), dim=1)
        flip = torch.rand(n, generator=g) < 0.1  # outliers the consistency pass must reject
        nrm[flip] = -nrm[flip]
        ids_l.append(ids); nrm_l.append(nrm); conf_l.append(torch.rand(n, generator=g) * 0.5 + 0.5)
    pcd = types.SimpleNamespace(_xyz=xyz)
    uid, sm = fusion(pcd, ids_l, nrm_l, conf_l, cams)
    out["fus_xyz"] = xyz.numpy(); out["fus_V"] = np.int32(V)
    for v in range(V):
        out[f"fus_E{v}"] = cams[v].extrinsics.numpy(); out[f"fus_ids{v}"] = ids_l[v].numpy()
        out[f"fus_n{v}"] = nrm_l[v].numpy(); out[f"fus_c{v}"] = conf_l[v].numpy()
    out["fus_unique_ids"] = uid.numpy(); out["fus_smoothed"] = sm.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
