"""The numpy restatement of the extraction post-pass (oracle/extract_oracle.py) against outputs of the reference's
own functions (tests/golden/extract_golden.npz, made by tests/golden/make_golden_extract.py with cv2 4.13)."""
import os

import numpy as np
import pytest

from oracle import extract_oracle as eo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "extract_golden.npz"))


@pytest.mark.parametrize("case", ["a", "b", "c", "s"])
def test_bilateral_matches_reference(case):
    kw = {}
    if case == "s":
        kw = dict(sigma_color=float(G["bil_s_sigma"][0]), sigma_space=float(G["bil_s_sigma"][1]))
    f, m = eo.masked_bilateral_filter(G[f"bil_{case}_depth"], G[f"bil_{case}_mask"], int(G[f"bil_{case}_d"]), **kw)
    assert np.array_equal(m, G[f"bil_{case}_newmask"].astype(bool))
    # OpenCV's range kernel is a 4096-bin interpolated table; tolerance = that table's error on a ~2 unit range
    np.testing.assert_allclose(f, G[f"bil_{case}_filtered"], rtol=0, atol=5e-6)
    assert np.abs(G[f"bil_{case}_filtered"] - G[f"bil_{case}_depth"]).max() > 0.05  # the filter did something


def test_bilateral_edge_cases():
    d = np.random.default_rng(0).random((9, 11)).astype(np.float32) + 1
    f, m = eo.masked_bilateral_filter(d, np.zeros_like(d, bool))  # nothing valid: depth passes through
    assert np.array_equal(f, d) and not m.any()
    f, m = eo.masked_bilateral_filter(d, np.ones_like(d, bool), d=1)  # 1x1 support: identity up to the normalisation
    assert m.all() and np.abs(f - d).max() < 1e-6


def fusion_inputs():
    V = int(G["fus_V"])
    return (G["fus_xyz"], [G[f"fus_ids{v}"] for v in range(V)], [G[f"fus_n{v}"] for v in range(V)],
            [G[f"fus_c{v}"] for v in range(V)], [G[f"fus_E{v}"][:3, 3] for v in range(V)])


def test_normal_fusion_matches_reference():
    uid, sm = eo.normal_fusion(*fusion_inputs())
    assert np.array_equal(uid, G["fus_unique_ids"])
    ref = G["fus_smoothed"]
    assert np.array_equal(np.isnan(sm), np.isnan(ref))  # Gaussians whose every observation was rejected: 0/0
    assert np.isnan(ref).sum() < ref.size / 2
    np.testing.assert_allclose(np.nan_to_num(sm), np.nan_to_num(ref), rtol=0, atol=2e-6)


def test_fusion_rejects_outliers():
    xyz, ids, nrm, conf, ts = fusion_inputs()
    uid, mean = eo.normal_fusion(xyz, ids, nrm, conf, ts, smooth=False)
    ok = ~np.isnan(mean).any(1)
    assert np.allclose(np.linalg.norm(mean[ok], axis=1), 1, atol=1e-5)
    assert len(uid) == len(np.unique(np.concatenate(ids)))


def test_bilateral_properties():
    """Size-independent properties of the filter (the same ones hold for the CUDA kernel; see test_gpu_extract.py)."""
    rng = np.random.default_rng(4)
    for H, W, d in [(17, 23, 3), (31, 19, 5), (8, 8, 7)]:
        depth = (1 + 4 * rng.random((H, W))).astype(np.float32)
        mask = rng.random((H, W)) > 0.15
        f, m = eo.masked_bilateral_filter(depth, mask, d=d, sigma_color=0.2, sigma_space=2.0)
        assert not (m & ~mask).any()                                  # the new mask only shrinks
        assert np.array_equal(f[~m], depth[~m])                       # masked-out pixels pass through bit for bit
        if m.sum() >= 2:
            lo, hi = depth[m].min(), depth[m].max()
            assert (f[m] >= lo - 1e-5).all() and (f[m] <= hi + 1e-5).all()   # a convex combination (0 == lo)
        # shrinking the mask can only shrink the result mask
        m2 = eo.masked_bilateral_filter(depth, mask & (rng.random((H, W)) > 0.1), d=d)[1]
        assert not (m2 & ~m).any()
    # an interior plateau far from invalid pixels and from other values is a fixed point
    depth = np.full((21, 21), 2.0, np.float32); depth[:, 12:] = 3.0
    f, m = eo.masked_bilateral_filter(depth, np.ones_like(depth, bool), d=3, sigma_color=1e-3, sigma_space=1.0)
    assert m.all() and np.abs(f - depth).max() < 1e-6                 # tiny range sigma: the edge is preserved
    # reference quirk: a zero depth range (constant depth, or a single surviving pixel) divides 0 by 0 -> NaN
    # (verified by running extract_pcd.py's function with cv2 4.13); restated, not "fixed"
    f, m = eo.masked_bilateral_filter(np.full((7, 9), 2.5, np.float32), np.ones((7, 9), bool))
    assert m.all() and np.isnan(f).all()
