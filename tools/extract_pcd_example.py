"""Example: the gs-extract-pcd flow of the reference (gaustudio/scripts/extract_pcd.py) on this stack, without meshing:
    python tools/extract_pcd_example.py --model point_cloud.ply --camera cameras.json [--sh 0] [-o fused.ply]
Loads a 3DGS PLY and cameras.json (gaustudio_b200.io), renders every view with the vanilla renderer, runs the GPU
post-pass (masked bilateral filter, normals, fusion) and writes the oriented surface point cloud.  With no arguments
it runs on a synthetic scene.  Not part of the test suite (needs a GPU); every piece it calls is."""
import argparse, sys
import torch
sys.path.insert(0, ".")
from gaustudio_b200 import extract, io, renderers
from gaustudio_b200.synthetic import build_config

ap = argparse.ArgumentParser()
ap.add_argument("--model", "-m"); ap.add_argument("--camera", "-c"); ap.add_argument("--sh", type=int, default=0)
ap.add_argument("--output", "-o", default="fused.ply")
a = ap.parse_args()
dev = torch.device("cuda")
if a.model:
    pcd = io.load_ply(a.model, active_sh_degree=a.sh, device=dev)
    cameras = io.load_cameras_json(a.camera)
else:
    pcd, cameras, _ = build_config("cfg1", P=50000, W=320, H=240, K=12)
    pcd.to(dev); pcd.active_sh_degree = a.sh
cameras = [c.to(dev) for c in cameras]
renderer = renderers.make({"name": "vanilla_renderer"})
xyz, rgb, normals, _ = extract.extract_pcd(renderer, pcd, cameras)
keep = ~torch.isnan(normals).any(1)
io.export_points_ply(a.output, xyz[keep], rgb[keep], normals[keep])
print(f"{int(keep.sum())} surface points from {len(cameras)} views -> {a.output}")
