"""numpy/ctypes front-end of the CPU oracle (oracle/gsr_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by gaustudio_b200/ (the product path).

Mirrors the call shape of the reference binding
(/root/reference/submodules/gaustudio-diff-gaussian-rasterization/rasterize_points.cu:35-210):
`forward(...)` -> dict(color, depth, median, opacity, radii, num_rendered), `backward(...)` -> dict of
the eight returned gradients (+ the two internal ones, dL_dconic / dL_ddepths).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    so = os.path.join(_HERE, "libgsr_oracle.so")
    src = os.path.join(_HERE, "gsr_oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgsr_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.gso_create_f32.restype = C.c_void_p
        _LIB.gso_create_f64.restype = C.c_void_p
        _LIB.gso_forward_f32.restype = C.c_int64
        _LIB.gso_forward_f64.restype = C.c_int64
        _LIB.gso_num_threads.restype = C.c_int
    return _LIB


def num_threads():
    return int(lib().gso_num_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One forward (+ optional backward) of the reference algorithm on the CPU."""

    def __init__(self, dtype=np.float32):
        self.dt = np.dtype(dtype)
        self.suf = "f32" if self.dt == np.float32 else "f64"
        self.real = C.c_float if self.dt == np.float32 else C.c_double
        self.h = C.c_void_p(getattr(lib(), "gso_create_" + self.suf)())

    def __del__(self):
        try:
            getattr(lib(), "gso_destroy_" + self.suf)(self.h)
        except Exception:
            pass

    def _a(self, x, shape=None):
        if x is None:
            return None
        x = np.ascontiguousarray(np.asarray(x, dtype=self.dt))
        if shape is not None:
            x = x.reshape(shape)
        return x

    def forward(self, means3D, opacities, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H, sh_degree=0,
                shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                scale_modifier=1.0):
        a = self._a
        self.means3D = a(means3D); P = self.means3D.shape[0]
        self.shs = a(shs); self.colors = a(colors_precomp); self.opac = a(opacities)
        self.scales = a(scales); self.rots = a(rotations); self.cov3D = a(cov3D_precomp)
        self.view = a(viewmatrix).reshape(16); self.proj = a(projmatrix).reshape(16); self.campos = a(campos)
        self.tan = (float(tanfovx), float(tanfovy)); self.mod = float(scale_modifier)
        self.W, self.H, self.P, self.D = int(W), int(H), P, int(sh_degree)
        self.M = 0 if self.shs is None else self.shs.shape[1]
        color = np.zeros((3, H, W), self.dt); depth = np.zeros((1, H, W), self.dt)
        median = np.zeros((3, H, W), self.dt); opacity = np.zeros((1, H, W), self.dt)
        radii = np.zeros(P, np.int32)
        R = 0
        if P:
            R = getattr(lib(), "gso_forward_" + self.suf)(
                self.h, P, self.D, self.M, self.W, self.H, _p(self.means3D), _p(self.shs), _p(self.colors),
                _p(self.opac), _p(self.scales), self.real(self.mod), _p(self.rots), _p(self.cov3D), _p(self.view),
                _p(self.proj), _p(self.campos), self.real(self.tan[0]), self.real(self.tan[1]), _p(color), _p(depth),
                _p(median), _p(opacity), _p(radii))
        self.R = int(R)
        return dict(color=color, depth=depth, median=median, opacity=opacity, radii=radii, num_rendered=int(R))

    def binning(self):
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        pl = np.zeros(self.R, np.uint32); rg = np.zeros((T, 2), np.uint32)
        nc = np.zeros((self.H, self.W), np.uint32); ft = np.zeros((self.H, self.W), self.dt)
        getattr(lib(), "gso_get_binning_" + self.suf)(self.h, _p(pl), _p(rg), _p(nc), _p(ft))
        return dict(point_list=pl, ranges=rg, n_contrib=nc, final_T=ft)

    def geometry(self):
        P = self.P
        g = dict(depths=np.zeros(P, self.dt), means2D=np.zeros((P, 2), self.dt),
                 conic_opacity=np.zeros((P, 4), self.dt), rgb=np.zeros((P, 3), self.dt),
                 cov3D=np.zeros((P, 6), self.dt), tiles_touched=np.zeros(P, np.uint32),
                 clamped=np.zeros((P, 3), np.uint8))
        getattr(lib(), "gso_get_geometry_" + self.suf)(self.h, _p(g["depths"]), _p(g["means2D"]),
                                                       _p(g["conic_opacity"]), _p(g["rgb"]), _p(g["cov3D"]),
                                                       _p(g["tiles_touched"]), _p(g["clamped"]))
        return g

    def backward(self, dL_color, dL_depth=None, dL_median=None, dL_opacity=None, bg=(0.0, 0.0, 0.0)):
        a = self._a
        P, M, H, W = self.P, self.M, self.H, self.W
        z = lambda *s: np.zeros(s, self.dt)
        dc = a(dL_color, (3, H, W))
        dd = a(dL_depth, (H, W)) if dL_depth is not None else z(H, W)
        dm = a(dL_median, (3, H, W)) if dL_median is not None else z(3, H, W)
        do = a(dL_opacity, (H, W)) if dL_opacity is not None else z(H, W)
        g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacities=z(P, 1), colors_precomp=z(P, 3), depths=z(P, 1),
                 means3D=z(P, 3), cov3D_precomp=z(P, 6), shs=z(P, M, 3), scales=z(P, 3), rotations=z(P, 4))
        if P:
            getattr(lib(), "gso_backward_" + self.suf)(
                self.h, _p(a(bg)), _p(self.means3D), _p(self.shs), _p(self.colors), _p(self.scales),
                self.real(self.mod), _p(self.rots), _p(self.cov3D), _p(self.view), _p(self.proj), _p(self.campos),
                self.real(self.tan[0]), self.real(self.tan[1]), _p(dc), _p(dd), _p(dm), _p(do), _p(g["means2D"]),
                _p(g["conic"]), _p(g["opacities"]), _p(g["colors_precomp"]), _p(g["depths"]), _p(g["means3D"]),
                _p(g["cov3D_precomp"]), _p(g["shs"]), _p(g["scales"]), _p(g["rotations"]))
        return g


def mark_visible(means3D, viewmatrix, dtype=np.float32):
    m = np.ascontiguousarray(means3D, dtype=dtype); v = np.ascontiguousarray(viewmatrix, dtype=dtype).reshape(16)
    out = np.zeros(m.shape[0], np.uint8)
    getattr(lib(), "gso_mark_visible_f32" if dtype == np.float32 else "gso_mark_visible_f64")(
        m.shape[0], _p(m), _p(v), _p(out))
    return out.astype(bool)


def depth2normal(depth, fx, fy, cx, cy, d_min=1e-3, d_max=1e5, rot=None, dtype=np.float32):
    d = np.ascontiguousarray(depth, dtype=dtype); H, W = d.shape
    out = np.zeros((H, W, 3), dtype)
    r = None if rot is None else np.ascontiguousarray(rot, dtype=dtype)
    real = C.c_float if dtype == np.float32 else C.c_double
    getattr(lib(), "gso_depth2normal_f32" if dtype == np.float32 else "gso_depth2normal_f64")(
        _p(d), W, H, real(fx), real(fy), real(cx), real(cy), real(d_min), real(d_max), _p(r), _p(out))
    return out
