"""On-disk formats either side of the path without `plyfile` / `omegaconf`:

* the 3DGS PLY point cloud as `BasePointCloud.load` reads it (`gaustudio/models/base.py:73-105`: one `vertex`
  element, float properties; attributes found by name prefix and ordered by their numeric suffix) and as
  `VanillaPointCloud.export` writes it (`gaustudio/models/vanilla_sg.py:144-181`: x y z nx ny nz f_dc_* f_rest_*
  opacity scale_* rot_*, little-endian float32, SH stored channel-major);
* `cameras.json` entries -> `Camera` (`gaustudio/utils/cameras_utils.py:8-38`).
"""
import json
import math
import re

import numpy as np
import torch

from .camera import Camera
from .synthetic import GaussianPointCloud

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def read_ply_vertices(path):
    """-> structured numpy array of the `vertex` element (ascii, binary_little_endian or binary_big_endian)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if props and not in_vertex:
                    pass
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
                elif count:  # an element after `vertex`: stop collecting properties
                    in_vertex = False
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties are not supported in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=count, ndmin=2)
            out = np.empty(count, dtype=[(n, t) for n, t in props])
            for i, (n, _) in enumerate(props):
                out[n] = data[:, i]
            return out
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, order + t) for n, t in props])
        return np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)


def _group(v, prefix):
    names = [n for n in v.dtype.names if n.startswith(prefix) and re.fullmatch(r".*_\d+", n)]
    names.sort(key=lambda n: int(n.split("_")[-1]))
    return np.stack([v[n] for n in names], axis=1).astype(np.float32) if names else None


def load_ply(path, sh_degree=3, active_sh_degree=0, device="cpu"):
    """PLY -> GaussianPointCloud holding the RAW attributes (log-scale, opacity logit, un-normalised rotation),
    `_f_dc` [P,3] and `_f_rest` [P,45] exactly as `BasePointCloud.load` stores them; `active_sh_degree` starts at
    0 like the reference model (`vanilla_sg.py:40`)."""
    v = read_ply_vertices(path)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1)
    f_dc, f_rest = _group(v, "f_dc"), _group(v, "f_rest")
    if f_rest is None:
        f_rest = np.zeros((len(v), 3 * ((sh_degree + 1) ** 2 - 1)), np.float32)
    return GaussianPointCloud(t(xyz), t(_group(v, "scale")), t(_group(v, "rot")), t(v["opacity"][:, None]), t(f_dc),
                              t(f_rest), sh_degree=sh_degree, active_sh_degree=active_sh_degree)


def export_ply(model, path):
    """GaussianPointCloud -> PLY in the reference's export layout (binary little-endian float32)."""
    n = len(model._xyz)
    flat = lambda a: a.detach().reshape(n, -1, 3).transpose(1, 2).flatten(start_dim=1).cpu().numpy()
    cols = [model._xyz.detach().cpu().numpy(), np.zeros((n, 3), np.float32), flat(model._f_dc), flat(model._f_rest),
            model._opacity.detach().reshape(n, 1).cpu().numpy(), model._scale.detach().cpu().numpy(),
            model._rot.detach().cpu().numpy()]
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(cols[2].shape[1])] + [f"f_rest_{i}" for i in range(cols[3].shape[1])]
    names += ["opacity"] + [f"scale_{i}" for i in range(cols[5].shape[1])] + [f"rot_{i}" for i in range(cols[6].shape[1])]
    data = np.concatenate(cols, axis=1).astype("<f4")
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
    header += "".join(f"property float {nm}\n" for nm in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(data.tobytes())


def export_points_ply(path, xyz, rgb=None, normals=None):
    """Surface point cloud (the `fused.ply` of gaustudio/scripts/extract_pcd.py:339-352, written there through
    open3d's `write_point_cloud`): binary little-endian, double x y z [nx ny nz] + uchar red green blue, the property
    layout open3d emits.  Tensors or arrays; colours in [0, 1]."""
    def host(a):
        return None if a is None else np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a)
    xyz, rgb, normals = host(xyz), host(rgb), host(normals)
    n = xyz.shape[0]
    fields = [("x", "<f8"), ("y", "<f8"), ("z", "<f8")]
    if normals is not None:
        fields += [("nx", "<f8"), ("ny", "<f8"), ("nz", "<f8")]
    if rgb is not None:
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.empty(n, dtype=fields)
    for i, k in enumerate("xyz"):
        rec[k] = xyz[:, i]
    if normals is not None:
        for i, k in enumerate(("nx", "ny", "nz")):
            rec[k] = normals[:, i]
    if rgb is not None:
        c = np.clip(np.nan_to_num(rgb.astype(np.float64)) * 255.0, 0, 255).astype(np.uint8)  # open3d truncates
        for i, k in enumerate(("red", "green", "blue")):
            rec[k] = c[:, i]
    names = {"<f8": "double", "u1": "uchar"}
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + \
           [f"property {names[t]} {k}" for k, t in fields] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        f.write(rec.tobytes())


def camera_from_json(entry):
    """One `cameras.json` entry (id, img_name, width, height, position, rotation (camera-to-world), fx, fy)."""
    c2w = np.eye(4)
    c2w[:3, :3] = np.array(entry["rotation"])
    c2w[:3, 3] = np.array(entry["position"])
    w2c = np.linalg.inv(c2w)
    fov = lambda focal, pixels: 2 * math.atan(pixels / (2 * focal))
    return Camera(R=w2c[:3, :3].transpose(), T=w2c[:3, 3], FoVx=fov(entry["fx"], entry["width"]),
                  FoVy=fov(entry["fy"], entry["height"]), image_width=entry["width"], image_height=entry["height"])


def load_cameras_json(path):
    with open(path) as f:
        return [camera_from_json(e) for e in json.load(f)]
