"""Generates tests/golden/camera_golden.npz by importing the reference's OWN Python classes on the CPU
(`gaustudio.datasets.Camera`, /root/reference/gaustudio/datasets/__init__.py:114-380).  Modules the reference
imports at package level but that are absent from this image (plyfile, ...) are stubbed -- none is touched by
Camera.  Run here (needs /root/reference):   python tests/golden/make_golden_camera.py
"""
import importlib
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return _Stub("call")


def _import_with_stubs(name):
    for _ in range(50):
        try:
            return importlib.import_module(name)
        except ModuleNotFoundError as e:
            top = e.name
            sys.modules[top] = _Stub(top)
    raise RuntimeError("too many missing modules")


ds = _import_with_stubs("gaustudio.datasets")
Camera = ds.Camera
rng = np.random.RandomState(5)
out = {}
for i in range(4):
    q = rng.randn(4); q /= np.linalg.norm(q)
    r, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                  [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                  [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
    T = rng.randn(3) * 2
    W, H = [(64, 48), (100, 60), (33, 47), (128, 72)][i]
    fovx, fovy = math.radians(40 + 10 * i), math.radians(35 + 7 * i)
    cam = Camera(R=R, T=T, FoVx=fovx, FoVy=fovy, image_width=W, image_height=H)
    depth = torch.tensor(rng.uniform(0.5, 4.0, (H, W)).astype(np.float32))
    depth[rng.rand(H, W) < 0.05] = 0.0  # invalid pixels
    out.update({f"c{i}_R": R, f"c{i}_T": T, f"c{i}_fov": np.array([fovx, fovy]), f"c{i}_wh": np.array([W, H]),
                f"c{i}_view": cam.world_view_transform.numpy(), f"c{i}_proj": cam.full_proj_transform.numpy(),
                f"c{i}_center": cam.camera_center.numpy(), f"c{i}_K": cam.intrinsics.numpy(),
                f"c{i}_depth": depth.numpy(), f"c{i}_normal_cam": cam.depth2normal(depth).numpy(),
                f"c{i}_normal_world": cam.depth2normal(depth, coordinate="world").numpy(),
                f"c{i}_point_cam": cam.depth2point(depth).numpy(),
                f"c{i}_point_world": cam.depth2point(depth, coordinate="world").numpy()})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "camera_golden.npz"), **out)
print("wrote camera_golden.npz", len(out), "arrays")
