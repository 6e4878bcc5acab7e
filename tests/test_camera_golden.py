"""Camera matrices + depth->normal against golden vectors produced by the reference's own Python classes
(tests/golden/make_golden_camera.py imports gaustudio.datasets.Camera from /root/reference)."""
import os

import numpy as np
import pytest

from gaustudio_b200.camera import Camera
from oracle import oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "camera_golden.npz"))


@pytest.mark.parametrize("i", range(4))
def test_camera_matrices_match_reference(i):
    W, H = [int(v) for v in G[f"c{i}_wh"]]
    cam = Camera(R=G[f"c{i}_R"], T=G[f"c{i}_T"], FoVx=float(G[f"c{i}_fov"][0]), FoVy=float(G[f"c{i}_fov"][1]),
                 image_width=W, image_height=H)
    np.testing.assert_array_equal(cam.world_view_transform.numpy(), G[f"c{i}_view"])
    np.testing.assert_allclose(cam.full_proj_transform.numpy(), G[f"c{i}_proj"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(cam.camera_center.numpy(), G[f"c{i}_center"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(cam.intrinsics.numpy(), G[f"c{i}_K"], rtol=1e-7)


@pytest.mark.parametrize("i", range(4))
@pytest.mark.parametrize("coord", ["cam", "world"])
def test_oracle_depth2normal_matches_reference(i, coord):
    K = G[f"c{i}_K"]
    rot = None
    if coord == "world":
        ext = G[f"c{i}_view"].T  # extrinsics = world_view_transform^T (datasets/__init__.py:219-221)
        rot = np.linalg.inv(ext[:3, :3].astype(np.float64)).T.astype(np.float32)
    n = orc.depth2normal(G[f"c{i}_depth"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], rot=rot)
    ref = G[f"c{i}_normal_{coord}"]
    assert n.shape == ref.shape
    invalid = (ref == -1).all(-1)
    assert ((n == -1).all(-1) == invalid).all()
    assert np.abs(n - ref).max() < 1e-4  # BASELINE tolerance for normals
    assert invalid.any() and (~invalid).any()
