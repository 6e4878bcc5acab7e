"""Generates tests/golden/ref_export.ply (+ ref_export_attrs.npz) by RUNNING the reference's own exporter and loader.

`VanillaPointCloud.export` / `construct_list_of_attributes` (gaustudio/models/vanilla_sg.py:144-181) and
`BasePointCloud.load` (gaustudio/models/base.py:73-105) are pulled out of the unmodified reference sources with `ast`
and executed on a 7-point model.  The reference goes through the `plyfile` package, which is not installed here; the
stand-in below implements only the container `plyfile` writes / reads for this call pattern (one `vertex` element of
float32 properties, `binary_little_endian 1.0`, header lines `property float <name>`).  What the fixture pins is the
reference's ATTRIBUTE PACKING -- names, order, the channel-major SH transposes, the suffix-sorted read-back -- i.e.
the part a loader gets wrong.  Nothing of the reference is written into the repo; only the produced file is.

    python tests/golden/make_golden_ply.py        (needs /root/reference; CPU only)
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF_SG = "/root/reference/gaustudio/models/vanilla_sg.py"
REF_BASE = "/root/reference/gaustudio/models/base.py"
OUT_PLY = os.path.join(HERE, "ref_export.ply")
OUT_NPZ = os.path.join(HERE, "ref_export_attrs.npz")


class PlyProperty:
    def __init__(self, name):
        self.name = name


class PlyElement:
    """plyfile.PlyElement for a structured float32 array."""

    def __init__(self, data, name):
        self.data, self.name, self.count = data, name, len(data)
        self.properties = [PlyProperty(n) for n in data.dtype.names]

    @staticmethod
    def describe(data, name):
        assert all(data.dtype[n] == np.dtype("f4") for n in data.dtype.names)
        return PlyElement(data, name)

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    """plyfile.PlyData: `PlyData([el]).write(path)` (binary, native = little endian) and `PlyData.read(path)`."""

    def __init__(self, elements):
        self.elements = list(elements)

    def __getitem__(self, name):
        return next(e for e in self.elements if e.name == name)

    def write(self, path):
        el = self.elements[0]
        head = ["ply", "format binary_little_endian 1.0", f"element {el.name} {el.count}"]
        head += [f"property float {n}" for n in el.data.dtype.names] + ["end_header"]
        with open(path, "wb") as f:
            f.write(("\n".join(head) + "\n").encode("ascii"))
            f.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply" and f.readline().split()[1] == b"binary_little_endian"
            _, name, count = f.readline().split()
            names = []
            while True:
                tok = f.readline().split()
                if tok[0] == b"end_header":
                    break
                assert tok[:2] == [b"property", b"float"]
                names.append(tok[2].decode())
            dt = np.dtype([(n, "<f4") for n in names])
            data = np.frombuffer(f.read(int(count) * dt.itemsize), dtype=dt)
        return PlyData([PlyElement(data, name.decode())])


def lift(path, cls, names):
    tree = ast.parse(open(path).read())
    c = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    fns = [n for n in c.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(fns) == len(names), (cls, names)
    ns = {"np": np, "torch": torch, "PlyData": PlyData, "PlyElement": PlyElement}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


class Shell:
    """Attribute holder the lifted methods run on (what they touch of `self`)."""
    device = "cpu"


def main():
    export, attrs = lift(REF_SG, "VanillaPointCloud", ["export", "construct_list_of_attributes"])
    (load,) = lift(REF_BASE, "BasePointCloud", ["load"])
    g = torch.Generator().manual_seed(21)
    P = 7
    m = Shell()
    m._xyz = torch.randn(P, 3, generator=g)
    m._f_dc = torch.randn(P, 1, 3, generator=g)
    m._f_rest = torch.randn(P, 15, 3, generator=g)
    m._opacity = torch.randn(P, 1, generator=g)
    m._scale = torch.randn(P, 3, generator=g)
    m._rot = torch.randn(P, 4, generator=g)
    m.construct_list_of_attributes = lambda: attrs(m)
    export(m, OUT_PLY)
    # read it back with the reference's own loader: what a gaustudio model holds after load()
    r = Shell()
    r.config = {"attributes": ["xyz", "opacity", "scale", "rot", "f_dc", "f_rest"]}
    load(r, OUT_PLY)
    np.savez(OUT_NPZ, **{k: getattr(m, "_" + k).numpy() for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rot")},
             **{"loaded_" + k: getattr(r, "_" + k).numpy() for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rot")})
    # and the other direction: THIS repo's exporter, given the same in-memory model, must write the same bytes -- and the
    # reference's loader must read them identically.  (export o load is NOT the identity in the reference: export writes
    # the SH channel-major, load keeps the 45 columns as they are and get_features reshapes them (P,15,3) without the
    # transpose back -- vanilla_sg.py:102-106 vs :147-148, SURVEY.md quirk 13.  Both directions are pinned as they are.)
    from gaustudio_b200 import io as gio
    from gaustudio_b200.synthetic import GaussianPointCloud
    ours = os.path.join(HERE, "_ours_tmp.ply")
    gio.export_ply(GaussianPointCloud(m._xyz, m._scale, m._rot, m._opacity, m._f_dc, m._f_rest), ours)
    assert open(ours, "rb").read() == open(OUT_PLY, "rb").read()
    r2 = Shell()
    r2.config = r.config
    load(r2, ours)
    for k in ("xyz", "f_dc", "f_rest", "opacity", "scale", "rot"):
        assert torch.equal(getattr(r, "_" + k), getattr(r2, "_" + k)), k
    os.remove(ours)
    print("wrote", OUT_PLY, os.path.getsize(OUT_PLY), "bytes;", OUT_NPZ)


if __name__ == "__main__":
    main()
