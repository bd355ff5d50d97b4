"""The oracle's numpy restatement of the geometric-consistency filter (incl. its own cv2.remap restatement) against
outputs of the reference's functions run with the real cv2 (tests/golden/fusion_filter.npz)."""
import numpy as np
import pytest

from oracle import damvs_oracle as O
from tests.golden_io import load_fusion_filter


@pytest.mark.parametrize("name", ["a", "b"])
def test_filter_matches_reference_fixture(name):
    fx = load_fusion_filter()
    K, E, d, c = fx[name + "/K"], fx[name + "/E"], fx[name + "/depths"], fx[name + "/confs"]
    masks, rep = O.check_geometric_consistency(d[0], K[0], E[0], d[1], K[1], E[1])
    want = fx[name + "/pair_masks"].astype(bool)
    assert (np.stack(masks) != want).mean() < 2e-4           # knife-edge pixels of the thresholds only
    both = masks[-1] & want[-1]
    np.testing.assert_allclose(rep[both], fx[name + "/pair_depth_reprojected"][both], rtol=1e-6, atol=1e-3)
    out = O.filter_reference_view(d[0], list(c), K[0], E[0], list(d[1:]), list(K[1:]), list(E[1:]))
    for k in ("photo_mask", "geo_mask", "final_mask"):
        assert (out[k] != fx[name + "/" + k].astype(bool)).mean() < 2e-4, k
    agree = out["geo_mask"] == fx[name + "/geo_mask"].astype(bool)
    err = np.abs(out["depth_est_averaged"] - fx[name + "/depth_est_averaged"])[agree]
    assert np.quantile(err, 0.999) < 1e-3 and (err > 1e-2).mean() < 2e-4


def test_remap_restatement_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(0)
    img = rs.rand(37, 53).astype(np.float32) * 100
    x = (rs.rand(40, 60) * 60 - 4).astype(np.float32)
    y = (rs.rand(40, 60) * 44 - 4).astype(np.float32)
    want = cv2.remap(img, x, y, interpolation=cv2.INTER_LINEAR)
    got = O.remap_bilinear(img, x, y)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-5)
