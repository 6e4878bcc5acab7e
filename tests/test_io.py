"""PLY / cameras.json I/O: layout of the reference's export, loader semantics of its load."""
import json
import math

import numpy as np
import torch

from gaustudio_b200 import io
from gaustudio_b200.camera import look_at_camera
from gaustudio_b200.synthetic import make_scene


def test_ply_roundtrip_and_layout(tmp_path):
    m = make_scene(257, 1.0, 0.05, seed=9)
    p = tmp_path / "pc.ply"
    io.export_ply(m, str(p))
    raw = p.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 257"]
    props = [l.split()[2] for l in lines if l.startswith("property")]
    # vanilla_sg.py:159-181 (construct_list_of_attributes)
    assert props == (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] +
                     [f"f_rest_{i}" for i in range(45)] + ["opacity"] + [f"scale_{i}" for i in range(3)] +
                     [f"rot_{i}" for i in range(4)])
    assert all(l.split()[1] == "float" for l in lines if l.startswith("property"))
    assert len(body) == 257 * len(props) * 4
    row0 = np.frombuffer(body[:len(props) * 4], "<f4")
    assert np.allclose(row0[:3], m._xyz[0].numpy()) and np.all(row0[3:6] == 0)
    # SH is written channel-major: f_rest_k = coefficient (k % 15)+1 of channel k // 15 (vanilla_sg.py:147-148)
    assert row0[6 + 3 + 16] == m._f_rest[0, 1, 1]
    back = io.load_ply(str(p))
    assert back.active_sh_degree == 0 and back.max_sh_degree == 3 and back.num_points == 257
    for name in ("_xyz", "_scale", "_rot", "_opacity"):
        assert torch.equal(getattr(back, name), getattr(m, name).reshape(getattr(back, name).shape))
    assert back._f_dc.shape == (257, 3) and back._f_rest.shape == (257, 45)
    # the loader keeps the file's flat order (base.py:94-104); get_features reshapes it as (P,-1,3) (vanilla_sg.py:102-106)
    assert torch.equal(back._f_rest, m._f_rest.transpose(1, 2).flatten(start_dim=1))
    assert back.get_features.shape == (257, 16, 3)
    assert torch.equal(back.get_features[:, 0], m._f_dc[:, 0])  # DC is layout-invariant ([P,1,3])


def test_ascii_ply_and_extra_elements(tmp_path):
    p = tmp_path / "a.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment hi\nelement vertex 2\nproperty float x\nproperty float y\n"
                 "property float z\nproperty float opacity\nproperty float scale_1\nproperty float scale_0\n"
                 "property float scale_2\nproperty float rot_0\nproperty float rot_1\nproperty float rot_2\n"
                 "property float rot_3\nproperty float f_dc_0\nproperty float f_dc_1\nproperty float f_dc_2\n"
                 "element face 0\nproperty list uchar int vertex_indices\nend_header\n"
                 "1 2 3 0.5 11 10 12 1 0 0 0 .1 .2 .3\n4 5 6 -0.5 21 20 22 0 1 0 0 .4 .5 .6\n")
    m = io.load_ply(str(p))
    assert m._xyz.tolist() == [[1, 2, 3], [4, 5, 6]]
    assert m._scale.tolist() == [[10, 11, 12], [20, 21, 22]]  # ordered by numeric suffix, not file order
    assert m._f_rest.shape == (2, 45) and float(m._f_rest.abs().max()) == 0


def test_cameras_json_matches_look_at(tmp_path):
    cam = look_at_camera((2.0, 1.0, 0.5), (0, 0, 0), 320, 240, math.radians(60), math.radians(47))
    c2w = np.linalg.inv(cam.world_view_transform.numpy().T.astype(np.float64))
    fx = 320 / (2 * math.tan(cam.FoVx / 2)); fy = 240 / (2 * math.tan(cam.FoVy / 2))
    entry = dict(id=0, img_name="v0", width=320, height=240, position=c2w[:3, 3].tolist(),
                 rotation=c2w[:3, :3].tolist(), fx=fx, fy=fy)
    p = tmp_path / "cameras.json"
    p.write_text(json.dumps([entry]))
    (back,) = io.load_cameras_json(str(p))
    assert np.allclose(back.world_view_transform.numpy(), cam.world_view_transform.numpy(), atol=1e-5)
    assert np.allclose(back.full_proj_transform.numpy(), cam.full_proj_transform.numpy(), atol=1e-5)
    assert abs(back.FoVx - cam.FoVx) < 1e-9 and (back.image_width, back.image_height) == (320, 240)


def test_surface_point_cloud_ply(tmp_path):
    g = np.random.default_rng(3)
    xyz = g.normal(size=(50, 3)).astype(np.float32); rgb = g.random((50, 3)).astype(np.float32)
    nrm = g.normal(size=(50, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    p = tmp_path / "fused.ply"
    io.export_points_ply(str(p), torch.from_numpy(xyz), torch.from_numpy(rgb), nrm)
    head = p.read_bytes().split(b"end_header\n")[0].decode().splitlines()
    assert head[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 50"]
    assert [l.split()[1:] for l in head[3:]] == [["double", k] for k in ("x", "y", "z", "nx", "ny", "nz")] + \
        [["uchar", k] for k in ("red", "green", "blue")]
    v = io.read_ply_vertices(str(p))
    assert np.allclose(np.stack([v["x"], v["y"], v["z"]], 1), xyz) and np.allclose(np.stack([v["nx"], v["ny"], v["nz"]], 1), nrm)
    assert np.array_equal(np.stack([v["red"], v["green"], v["blue"]], 1), (rgb.astype(np.float64) * 255).astype(np.uint8))
    io.export_points_ply(str(p), xyz)  # positions only
    assert io.read_ply_vertices(str(p)).dtype.names == ("x", "y", "z")


def test_ply_written_by_the_reference_exporter():
    """tests/golden/ref_export.ply was produced by the reference's own `VanillaPointCloud.export` (executed unmodified,
    tests/golden/make_golden_ply.py); `loaded_*` is what the reference's own `BasePointCloud.load` reads back from it.
    Our loader must hold the same tensors, and our exporter must write the same bytes for the same in-memory model."""
    import os
    import tempfile
    import numpy as np
    import torch
    from gaustudio_b200 import io as gio
    from gaustudio_b200.synthetic import GaussianPointCloud
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ply, G = os.path.join(here, "ref_export.ply"), np.load(os.path.join(here, "ref_export_attrs.npz"))
    m = gio.load_ply(ply)
    for k in ("xyz", "opacity", "scale", "rot", "f_dc", "f_rest"):
        ours = getattr(m, "_" + k).numpy()
        assert np.array_equal(ours.reshape(G["loaded_" + k].shape), G["loaded_" + k]), k
    # the SH tensor the rasterizer consumes, built exactly as the reference model builds it from what it loaded
    # (vanilla_sg.py:102-106: reshape(P, -1, 3) of the loaded [P,3] / [P,45] arrays)
    P = G["xyz"].shape[0]
    want = np.concatenate([G["loaded_f_dc"].reshape(P, -1, 3), G["loaded_f_rest"].reshape(P, -1, 3)], axis=1)
    assert np.array_equal(m.get_features.numpy(), want)
    # our exporter on the in-memory model the reference exported (export o load is not the identity in the reference:
    # SH goes out channel-major and comes back un-transposed, vanilla_sg.py:102-106 vs :147-148)
    t = lambda k: torch.from_numpy(G[k])
    src = GaussianPointCloud(t("xyz"), t("scale"), t("rot"), t("opacity"), t("f_dc"), t("f_rest"))
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.ply")
        gio.export_ply(src, out)
        assert open(out, "rb").read() == open(ply, "rb").read()
