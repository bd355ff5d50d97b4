"""Golden fixture for the geometric-consistency filter, produced by the REFERENCE's own
functions: ``reproject_with_depth`` / ``check_geometric_consistency`` (filter/dypcd.py:98-159) are extracted from the
unmodified file (the module itself imports plyfile, which is not installed) and executed with the real ``cv2`` and
the reference's default thresholds (test_uni.py:104-109); the accumulation over source views is the loop of
filter_depth (filter/dypcd.py:224-257) transcribed around those calls.
Runs only in the build container.   Usage: python tests/golden/make_golden_filter.py  (writes fusion_filter.npz)
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def load_reference_functions():
    import cv2
    path = "/root/reference/filter/dypcd.py"
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("reproject_with_depth", "check_geometric_consistency")]
    assert len(keep) == 2
    mod = types.ModuleType("ref_dypcd_subset")
    mod.__dict__.update({"np": np, "cv2": cv2})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), mod.__dict__)
    return mod


def make_scene(seed, H, W, n_src):
    """A smooth surface seen by n_src + 1 nearby cameras; every view's depth map is rendered from the surface by a
    fixed-point iteration and perturbed, so that most pixels are consistent and some are not."""
    rs = np.random.RandomState(seed)
    f = 0.9 * W
    K = np.array([[f, 0, W / 2 - 0.5], [0, f, H / 2 - 0.5], [0, 0, 1]], np.float32)

    def surface(xw, yw):
        return 600 + 40 * np.sin(xw / 90.0) * np.cos(yw / 70.0)

    def rot(a, b, c):
        ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
        return (np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]]) @ np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]]) @
                np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]]))
    Es, Ks, depths = [], [], []
    for v in range(n_src + 1):
        E = np.eye(4)
        if v:
            E[:3, :3] = rot(*rs.uniform(-0.03, 0.03, 3))
            E[:3, 3] = rs.uniform(-40, 40, 3)
        Kv = K.copy()
        Kv[0, 0] *= 1 + 0.02 * v
        Kv[1, 1] *= 1 + 0.02 * v
        x, y = np.meshgrid(np.arange(W), np.arange(H))
        d = np.full((H, W), 600.0)
        Ei, Ki = np.linalg.inv(E), np.linalg.inv(Kv.astype(np.float64))
        for _ in range(30):   # depth along each ray such that the world point lies on z_world = surface(x_world, y_world)
            cam = Ki @ (np.vstack((x.reshape(-1), y.reshape(-1), np.ones(H * W))) * d.reshape(-1))
            world = Ei[:3, :3] @ cam + Ei[:3, 3:4]
            err = surface(world[0], world[1]) - world[2]
            d = d + (err / Ei[2, :3].dot(Ki[:, 2])).reshape(H, W) * 0.9
        noise = rs.normal(0, 0.15, (H, W)) + (rs.rand(H, W) < 0.05) * rs.normal(0, 15, (H, W))
        depths.append((d + noise).astype(np.float32))
        Es.append(E.astype(np.float32))
        Ks.append(Kv)
    return Ks, Es, depths


def main():
    ref = load_reference_functions()
    args = types.SimpleNamespace(dist_base=1 / 4, rel_diff_base=1 / 1300, conf=[0.1, 0.15, 0.9])
    blob = {}
    for name, (H, W, n_src, seed) in {"a": (96, 128, 4, 1), "b": (80, 112, 9, 2)}.items():
        Ks, Es, depths = make_scene(seed, H, W, n_src)
        rs = np.random.RandomState(seed + 100)
        confs = [rs.rand(H, W).astype(np.float32) * 0.5 + 0.05, rs.rand(H, W).astype(np.float32) * 0.6 + 0.1,
                 rs.rand(H, W).astype(np.float32) * 0.3 + 0.75]
        ref_depth = depths[0]
        photo_mask = np.logical_and(np.logical_and(confs[2] > args.conf[2], confs[1] > args.conf[1]), confs[0] > args.conf[0])
        all_depth, geo_mask_sum = [], 0
        dy_range = n_src + 1
        geo_mask_sums = [0] * (dy_range - 2)
        for v in range(1, n_src + 1):
            masks, geo_mask, depth_reprojected, x2d_src, y2d_src = ref.check_geometric_consistency(
                args, ref_depth, Ks[0], Es[0], depths[v], Ks[v], Es[v])
            geo_mask_sum += geo_mask.astype(np.int32)
            for i in range(2, dy_range):
                geo_mask_sums[i - 2] += masks[i - 2].astype(np.int32)
            all_depth.append(depth_reprojected)
            if v == 1:
                blob[name + "/pair_masks"] = np.stack(masks).astype(np.uint8)
                blob[name + "/pair_depth_reprojected"] = depth_reprojected
        depth_est_averaged = (sum(all_depth) + ref_depth) / (geo_mask_sum + 1)
        geo_mask = geo_mask_sum >= dy_range
        for i in range(2, dy_range):
            geo_mask = np.logical_or(geo_mask, geo_mask_sums[i - 2] >= i)
        final_mask = np.logical_and(photo_mask, geo_mask)
        blob[name + "/K"] = np.stack(Ks)
        blob[name + "/E"] = np.stack(Es)
        blob[name + "/depths"] = np.stack(depths)
        blob[name + "/confs"] = np.stack(confs)
        blob[name + "/depth_est_averaged"] = depth_est_averaged.astype(np.float32)
        blob[name + "/photo_mask"] = photo_mask.astype(np.uint8)
        blob[name + "/geo_mask"] = geo_mask.astype(np.uint8)
        blob[name + "/final_mask"] = final_mask.astype(np.uint8)
        print(name, "geo %.3f photo %.3f final %.3f" % (geo_mask.mean(), photo_mask.mean(), final_mask.mean()))
    np.savez_compressed(os.path.join(HERE, "fusion_filter.npz"), **blob)
    print("done", os.path.getsize(os.path.join(HERE, "fusion_filter.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
