"""GPU parity of the fused geometric-consistency filter (damvs_geo_consistency_fuse through the C ABI) against the
reference-generated fixture and the numpy oracle.  Masks are thresholded comparisons: they must agree everywhere
except on knife-edge pixels (< 2e-4 of the image); the averaged depth within 1e-3 where the masks agree."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import damvs_oracle as O  # noqa: E402
from tests.golden_io import load_fusion_filter  # noqa: E402


def dev():
    return torch.device("cuda:0")


def run_native(d, c, K, E, **kw):
    from damvsnet_b200 import fusion
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    out = fusion.filter_reference_view(t(d[0]), [t(x) for x in c], K[0], E[0], [t(x) for x in d[1:]], list(K[1:]), list(E[1:]), **kw)
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("name", ["a", "b"])
def test_filter_matches_reference_fixture(name):
    fx = load_fusion_filter()
    K, E, d, c = fx[name + "/K"], fx[name + "/E"], fx[name + "/depths"], fx[name + "/confs"]
    got = run_native(d, c, K, E)
    for k in ("photo_mask", "geo_mask", "final_mask"):
        assert (got[k] != fx[name + "/" + k].astype(bool)).mean() < 2e-4, k
    agree = got["geo_mask"] == fx[name + "/geo_mask"].astype(bool)
    err = np.abs(got["depth_est_averaged"] - fx[name + "/depth_est_averaged"])[agree]
    assert np.quantile(err, 0.999) < 1e-3 and (err > 1e-2).mean() < 2e-4


def test_filter_matches_oracle_other_thresholds_and_identity():
    fx = load_fusion_filter()
    K, E, d, c = fx["a/K"], fx["a/E"], fx["a/depths"], fx["a/confs"]
    kw = dict(conf_thr=(0.2, 0.3, 0.8), dist_base=0.4, rel_diff_base=1 / 900)
    got = run_native(d, c, K, E, **kw)
    want = O.filter_reference_view(d[0], list(c), K[0], E[0], list(d[1:]), list(K[1:]), list(E[1:]), **kw)
    for k in ("photo_mask", "geo_mask", "final_mask"):
        assert (got[k] != want[k]).mean() < 2e-4, k
    # a view checked against copies of itself is consistent everywhere and averages to itself
    same = run_native(np.stack([d[0]] * 3), c, np.stack([K[0]] * 3), np.stack([E[0]] * 3))
    assert same["geo_mask"].all()
    np.testing.assert_allclose(same["depth_est_averaged"], d[0], rtol=1e-5)


def test_backproject_valid_points():
    from damvsnet_b200 import fusion
    fx = load_fusion_filter()
    K, E, d = fx["a/K"], fx["a/E"], fx["a/depths"]
    mask = torch.from_numpy(fx["a/final_mask"].astype(bool)).to(dev())
    pts = fusion.backproject_valid(torch.from_numpy(d[1]).to(dev()), mask, K[1], E[1]).cpu().numpy()
    ys, xs = np.nonzero(fx["a/final_mask"])
    xyz = np.matmul(np.linalg.inv(K[1]), np.vstack((xs, ys, np.ones_like(xs))) * d[1][ys, xs])
    world = np.matmul(np.linalg.inv(E[1]), np.vstack((xyz, np.ones_like(xs))))[:3].T          # filter/dypcd.py:294-298
    np.testing.assert_allclose(pts, world, rtol=1e-5, atol=1e-3)


def test_fuse_views_and_ply_round_trip(tmp_path):
    """A three-view scan fused on the GPU: the points of every reference view equal the reference's back-projection of
    the oracle's averaged depth under the oracle's final mask, and the PLY file parses back to the same vertices."""
    from damvsnet_b200 import fusion
    fx = load_fusion_filter()
    K, E, d, c = fx["a/K"][:3], fx["a/E"][:3], fx["a/depths"][:3], fx["a/confs"]
    rs = np.random.RandomState(5)
    imgs = [rs.rand(*d[0].shape, 3).astype(np.float32) for _ in range(3)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    pairs = [(0, [1, 2]), (1, [0, 2]), (2, [1, 0])]
    xyz, rgb = fusion.fuse_views([t(x) for x in d], [[t(x) for x in c]] * 3, [t(x) for x in imgs], list(K), list(E), pairs)
    want_xyz, want_rgb = [], []
    for ref, srcs in pairs:
        o = O.filter_reference_view(d[ref], list(c), K[ref], E[ref], [d[s] for s in srcs], [K[s] for s in srcs], [E[s] for s in srcs])
        ys, xs = np.nonzero(o["final_mask"])
        dep = o["depth_est_averaged"][ys, xs]
        cam = np.matmul(np.linalg.inv(K[ref]), np.vstack((xs, ys, np.ones_like(xs))) * dep)
        want_xyz.append(np.matmul(np.linalg.inv(E[ref]), np.vstack((cam, np.ones_like(xs))))[:3].T)
        want_rgb.append((imgs[ref][o["final_mask"]] * 255).astype(np.uint8))
    want_xyz, want_rgb = np.concatenate(want_xyz), np.concatenate(want_rgb)
    assert abs(len(xyz) - len(want_xyz)) <= 3                      # knife-edge pixels of the masks
    if len(xyz) == len(want_xyz):
        np.testing.assert_allclose(xyz.cpu().numpy(), want_xyz, rtol=1e-5, atol=2e-2)
        assert (rgb.cpu().numpy() == want_rgb).mean() > 0.999
    path = str(tmp_path / "scan.ply")
    fusion.write_ply(path, xyz, rgb)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex %d" % len(xyz) in head and b"property uchar blue" in head
    vert = np.frombuffer(body, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    assert len(vert) == len(xyz)
    np.testing.assert_array_equal(vert["z"], xyz[:, 2].cpu().numpy())
    np.testing.assert_array_equal(vert["green"], rgb[:, 1].cpu().numpy())


def test_full_size_against_oracle():
    """DTU-test size (1152x1600, 4 source views): the masks agree with the numpy oracle on (almost) every pixel; the
    oracle's wall time -- the reference's single-core CPU algorithm -- is printed next to the kernel's (pytest -s)."""
    import os, sys, time
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_filter import make_scene
    H, W, n = 1152, 1600, 4
    Ks, Es, depths = make_scene(3, H, W, n)
    rs = np.random.RandomState(0)
    confs = [rs.rand(H, W).astype(np.float32) for _ in range(3)]
    t0 = time.perf_counter()
    want = O.filter_reference_view(depths[0], confs, Ks[0], Es[0], depths[1:], Ks[1:], Es[1:])
    cpu_s = time.perf_counter() - t0
    got = run_native(np.stack(depths), confs, np.stack(Ks), np.stack(Es))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_native(np.stack(depths), confs, np.stack(Ks), np.stack(Es))
    gpu_s = time.perf_counter() - t0        # includes the host->device copies of this helper
    print(f"filter 1152x1600 x4: numpy oracle {cpu_s:.2f} s, native call incl. uploads {gpu_s * 1e3:.1f} ms")
    for k in ("photo_mask", "geo_mask", "final_mask"):
        assert (got[k] != want[k]).mean() < 1e-5, k
