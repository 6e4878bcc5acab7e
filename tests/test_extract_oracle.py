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
