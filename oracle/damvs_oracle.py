"""CPU oracle for the DA-MVSNet per-stage cost-volume hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``damvsnet_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may.  The product path is the CUDA extension and it
fails loudly when that extension is missing.

This is a restatement, in plain fp32 torch-CPU tensor arithmetic, of what the
reference computes in ``DepthNet.forward`` (reference models/cas_mvsnet.py:18-134)
and the functions it calls.  Each function cites the reference lines it follows.
The arithmetic of the reference lives in PyTorch library ops (``F.grid_sample``,
``nn.Conv3d``, ``nn.BatchNorm3d``, ``F.softmax`` ...); the warp, the view-weight
network, the BatchNorm fold and the regression head are restated here in closed
form (no ``grid_sample``, no ``nn.Module``), the 3-D convolutions call
``F.conv3d`` / ``F.conv_transpose3d`` directly.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run
in the build container by ``tests/golden/make_golden.py`` (which imports the
unmodified reference from /root/reference) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against
those fixtures on every CPU test run.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm3d default, reference models/module.py:141,184


# --------------------------------------------------------------------------
# projection + warp
# --------------------------------------------------------------------------
def compose_projection(proj_pair: torch.Tensor) -> torch.Tensor:
    """[B,2,4,4] (extrinsic, intrinsic) -> [B,4,4] with rows 0-2 = K @ E[:3,:4].

    Reference models/cas_mvsnet.py:44-47.
    """
    out = proj_pair[:, 0].clone()
    out[:, :3, :4] = torch.matmul(proj_pair[:, 1, :3, :3], proj_pair[:, 0, :3, :4])
    return out


def relative_projection(src_proj: torch.Tensor, ref_proj: torch.Tensor):
    """rot [B,3,3], trans [B,3] of ``src_proj @ inverse(ref_proj)``.

    Reference models/module.py:308-310.
    """
    proj = torch.matmul(src_proj, torch.inverse(ref_proj))
    return proj[:, :3, :3].contiguous(), proj[:, :3, 3].contiguous()


def warp_coordinates(rot: torch.Tensor, trans: torch.Tensor, depth_values: torch.Tensor,
                     height: int, width: int):
    """Un-normalised source sample coordinates (ix, iy), each [B,D,H*W].

    Reference models/module.py:312-325 builds a grid normalised by (W-1)/2 and
    (H-1)/2; ``F.grid_sample`` is then called with its default
    ``align_corners=False`` (models/module.py:328-329), whose un-normalisation is
    ``((g + 1) * size - 1) / 2``.  Both steps are kept, in that order, so the
    rounding matches.
    """
    b = rot.shape[0]
    d = depth_values.shape[1]
    y, x = torch.meshgrid(torch.arange(0, height, dtype=torch.float32),
                          torch.arange(0, width, dtype=torch.float32), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(height * width)))  # [3,HW]
    rot_xyz = torch.matmul(rot, xyz.unsqueeze(0).expand(b, -1, -1))                 # [B,3,HW]
    dv = depth_values.reshape(b, 1, d, -1)                                          # [B,1,D,HW or 1]
    proj_xyz = rot_xyz.unsqueeze(2) * dv + trans.view(b, 3, 1, 1)                   # [B,3,D,HW]
    px = proj_xyz[:, 0] / proj_xyz[:, 2]
    py = proj_xyz[:, 1] / proj_xyz[:, 2]
    gx = px / ((width - 1) / 2) - 1
    gy = py / ((height - 1) / 2) - 1
    ix = ((gx + 1) * width - 1) / 2
    iy = ((gy + 1) * height - 1) / 2
    return ix, iy


def bilinear_zeros(src: torch.Tensor, ix: torch.Tensor, iy: torch.Tensor) -> torch.Tensor:
    """Bilinear sample with zero padding per tap.  src [B,C,H,W]; ix, iy [B,P].

    Restates ATen's grid_sampler_2d (bilinear, padding_mode='zeros'): taps at
    floor and floor+1, weights from the unclipped coordinate, an out-of-range
    tap contributes zero.
    """
    b, c, h, w = src.shape
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    wx1 = ix - x0
    wy1 = iy - y0
    wx0 = 1.0 - wx1
    wy0 = 1.0 - wy1
    flat = src.reshape(b, c, h * w)
    out = torch.zeros(b, c, ix.shape[1], dtype=src.dtype)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = x0 + dx
            yi = y0 + dy
            ok = (xi >= 0) & (xi <= w - 1) & (yi >= 0) & (yi <= h - 1)
            lin = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).long()
            tap = torch.gather(flat, 2, lin.unsqueeze(1).expand(-1, c, -1))
            out = out + tap * (wy * wx * ok.to(src.dtype)).unsqueeze(1)
    return out


def homo_warping(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor,
                 depth_values: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] source features -> [B,C,D,H,W] warped volume.

    Reference models/module.py:297-332.  ``depth_values`` is [B,D] or [B,D,H,W].
    """
    b, c, h, w = src_fea.shape
    d = depth_values.shape[1]
    rot, trans = relative_projection(src_proj, ref_proj)
    ix, iy = warp_coordinates(rot, trans, depth_values, h, w)
    ix = ix.expand(b, d, h * w).reshape(b, -1)
    iy = iy.expand(b, d, h * w).reshape(b, -1)
    return bilinear_zeros(src_fea, ix, iy).view(b, c, d, h, w)


# --------------------------------------------------------------------------
# BatchNorm fold, view-weight net, aggregation
# --------------------------------------------------------------------------
def bn_affine(sd: Dict[str, torch.Tensor], prefix: str):
    """Eval-mode BatchNorm as y = x * scale + shift (reference models/module.py:141,150)."""
    scale = sd[prefix + ".weight"] / torch.sqrt(sd[prefix + ".running_var"] + BN_EPS)
    shift = sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * scale
    return scale, shift


def batch_norm(y: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, training: bool) -> torch.Tensor:
    """nn.BatchNorm3d(momentum=0.1, eps=1e-5) (reference models/module.py:141,184).  Eval: the folded affine.
    Training: normalise with the biased batch variance over (B,D,H,W); the running buffers are not touched
    here (the oracle is functional; buffer updates are checked separately in the tests)."""
    if training:
        return F.batch_norm(y, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.1, BN_EPS)
    s, b = bn_affine(sd, prefix)
    shape = (1, -1) + (1,) * (y.dim() - 2)
    return y * s.view(shape) + b.view(shape)


def view_weight(sq: torch.Tensor, sd: Dict[str, torch.Tensor], stage_idx: int, training: bool = False) -> torch.Tensor:
    """AggWeightNetVolume.forward: [B,C,D,H,W] -> [B,1,D,H,W].

    Reference models/module.py:544-563: two 1x1x1 Conv3d(bias=False)+BN+ReLU
    blocks (``w_net``); ``conv0`` is constructed but never called.
    """
    p = f"DepthNet.weight_net.{stage_idx}.w_net."
    w1 = sd[p + "0.conv.weight"].view(1, -1, 1, 1, 1)
    w2 = sd[p + "1.conv.weight"].view(())
    s = (sq * w1).sum(dim=1, keepdim=True)
    a = torch.relu(batch_norm(s, sd, p + "0.bn", training))
    return torch.relu(batch_norm(a * w2, sd, p + "1.bn", training))


def aggregate(features: Sequence[torch.Tensor], proj_matrices: torch.Tensor, depth_values: torch.Tensor,
              mode: str, sd: Optional[Dict[str, torch.Tensor]] = None, stage_idx: int = 0,
              training: bool = False) -> torch.Tensor:
    """Multi-view cost volume [B,C,D,H,W].  Reference models/cas_mvsnet.py:19-87."""
    projs = torch.unbind(proj_matrices, 1)
    assert len(features) == len(projs)
    n = len(features)
    ref, srcs = features[0], features[1:]
    ref_proj = compose_projection(projs[0])
    d = depth_values.shape[1]
    ref_vol = ref.unsqueeze(2).expand(-1, -1, d, -1, -1)
    if mode == "variance":
        vsum = ref_vol.clone()
        vsq = ref_vol ** 2
        for src, pm in zip(srcs, projs[1:]):
            wv = homo_warping(src, compose_projection(pm), ref_proj, depth_values)
            vsum = vsum + wv
            vsq = vsq + wv ** 2
        return vsq / n - (vsum / n) ** 2                      # cas_mvsnet.py:85
    if mode == "adaptive":
        acc = None
        for src, pm in zip(srcs, projs[1:]):
            wv = homo_warping(src, compose_projection(pm), ref_proj, depth_values)
            sq = (ref_vol - wv) ** 2                          # cas_mvsnet.py:66
            wt = view_weight(sq, sd, stage_idx, training)     # cas_mvsnet.py:71
            term = (wt + 1) * sq
            acc = term if acc is None else acc + term         # cas_mvsnet.py:73-76
        return acc / (n - 1)                                  # cas_mvsnet.py:87
    raise ValueError(mode)


# --------------------------------------------------------------------------
# CostRegNet
# --------------------------------------------------------------------------
def conv_block(x, sd, prefix, stride=1, training=False):
    """Conv3d(k3, padding=1, bias=False) + BN + ReLU.  Reference models/module.py:117-159."""
    y = F.conv3d(x, sd[prefix + ".conv.weight"], None, stride=stride, padding=1)
    return torch.relu(batch_norm(y, sd, prefix + ".bn", training))


def deconv_block(x, sd, prefix, training=False):
    """ConvTranspose3d(k3, s2, p1, op1, bias=False) + BN + ReLU.  Reference models/module.py:161-202."""
    y = F.conv_transpose3d(x, sd[prefix + ".conv.weight"], None, stride=2, padding=1, output_padding=1)
    return torch.relu(batch_norm(y, sd, prefix + ".bn", training))


def cost_reg_net(x: torch.Tensor, sd: Dict[str, torch.Tensor], stage_idx: int, training: bool = False) -> torch.Tensor:
    """[B,C,D,H,W] -> logits [B,1,D,H,W].  Reference models/module.py:510-541."""
    p = f"cost_regularization.{stage_idx}."
    t = training
    c0 = conv_block(x, sd, p + "conv0", 1, t)
    c2 = conv_block(conv_block(c0, sd, p + "conv1", 2, t), sd, p + "conv2", 1, t)
    c4 = conv_block(conv_block(c2, sd, p + "conv3", 2, t), sd, p + "conv4", 1, t)
    y = conv_block(conv_block(c4, sd, p + "conv5", 2, t), sd, p + "conv6", 1, t)
    y = c4 + deconv_block(y, sd, p + "conv7", t)
    y = c2 + deconv_block(y, sd, p + "conv9", t)
    y = c0 + deconv_block(y, sd, p + "conv11", t)
    return F.conv3d(y, sd[p + "prob.weight"], None, stride=1, padding=1)


# --------------------------------------------------------------------------
# regression head
# --------------------------------------------------------------------------
def regress_head(logits: torch.Tensor, depth_values: torch.Tensor) -> Dict[str, torch.Tensor]:
    """logits [B,D,H,W], depth_values [B,D,H,W] -> the five per-stage outputs.

    Reference models/cas_mvsnet.py:105-134 and models/module.py:609-615.
    """
    b, d, h, w = logits.shape
    m = logits.max(dim=1, keepdim=True).values
    e = torch.exp(logits - m)
    prob = e / e.sum(dim=1, keepdim=True)                                   # F.softmax(dim=1)
    depth = (prob * depth_values).sum(dim=1)                                # depth_regression
    # confidence: sum of p over [idx-1, idx+2], zero outside [0,D-1] (cas_mvsnet.py:113-118)
    padded = F.pad(prob, (0, 0, 0, 0, 1, 2))
    sum4 = padded[:, 0:d] + padded[:, 1:d + 1] + padded[:, 2:d + 2] + padded[:, 3:d + 3]
    k = torch.arange(d, dtype=torch.float32).view(1, d, 1, 1)
    idx = (prob * k).sum(dim=1).long().clamp(0, d - 1)
    conf = torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)
    var = 3 * torch.sum((depth_values - depth.unsqueeze(1)) ** 2 * prob, dim=1) ** 0.5   # cas_mvsnet.py:121-124
    return {"depth": depth, "photometric_confidence": conf, "variance": var,
            "prob_volume": prob, "depth_values": depth_values}


def depth_regression(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """Reference models/module.py:609-615."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(p * depth_values, 1)


# --------------------------------------------------------------------------
# depth-hypothesis sampling (upstream neighbour of the path)
# --------------------------------------------------------------------------
HYP_EPS = 1e-12  # reference models/module.py:10


def uncertainty_aware_samples(cur_depth: torch.Tensor, exp_var: torch.Tensor, ndepth: int) -> torch.Tensor:
    """cur_depth, exp_var [B,1,H,W] -> hypotheses [B,D,H,W].  Reference models/module.py:1012-1036 (the
    `cur_depth.dim() != 2` branch), without the Python lists: the D planes are one broadcast."""
    low = -torch.min(cur_depth, exp_var)
    step = (exp_var - low) / (float(ndepth) - 1)
    i = torch.arange(ndepth, dtype=cur_depth.dtype).view(1, ndepth, 1, 1)
    lin = low + step * i
    offset = F.softmax(3 * lin / (exp_var + HYP_EPS), dim=1)
    return cur_depth + lin + HYP_EPS + offset * step


def first_stage_samples(depth_values: torch.Tensor, ndepth: int) -> torch.Tensor:
    """depth_values [B,Dtot] -> [B,D] evenly spaced over its range.  Reference models/module.py:1003-1010
    (spatially constant, so the repeat and the resample of the caller are the identity on every pixel)."""
    lo, hi = depth_values[:, 0], depth_values[:, -1]
    interval = (hi - lo) / (ndepth - 1)
    return lo.unsqueeze(1) + torch.arange(ndepth, dtype=depth_values.dtype).view(1, -1) * interval.unsqueeze(1)


def stage_hypotheses(prev_depth: torch.Tensor, prev_var: torch.Tensor, ndepth: int, height: int, width: int,
                     scale: int) -> torch.Tensor:
    """[B,hp,wp] depth / variance of the previous stage -> depth_values [B,D,height/scale,width/scale] handed to
    DepthNet.forward.  Reference models/cas_mvsnet.py:250-253 (bilinear to full resolution), :269-274,
    :293-296 (trilinear to the stage resolution), all with align_corners=False."""
    cur = F.interpolate(prev_depth.unsqueeze(1), [height, width], mode="bilinear", align_corners=False)
    ev = F.interpolate(prev_var.unsqueeze(1), [height, width], mode="bilinear", align_corners=False)
    full = uncertainty_aware_samples(cur, ev, ndepth)
    return F.interpolate(full.unsqueeze(1), [ndepth, height // scale, width // scale], mode="trilinear",
                         align_corners=False).squeeze(1)


# --------------------------------------------------------------------------
# geometric-consistency filtering (downstream neighbour of the path)
# --------------------------------------------------------------------------
def remap_bilinear(img, x, y):
    """cv2.remap(img, x, y, interpolation=cv2.INTER_LINEAR) for float32 single-channel images and float32 maps
    (default BORDER_CONSTANT, value 0), restated in numpy from OpenCV's published algorithm (the dependency is
    third party: opencv-python 4.x, modules/imgproc/src/imgwarp.cpp remapBilinear; the reference calls it at
    filter/dypcd.py:118): map coordinates are rounded to 1/32 pixel (cvRound = round half to even), the four
    weights are float32 products of the 1-D table entries, taps outside the image contribute 0."""
    import numpy as np
    h, w = img.shape
    sx = np.rint(x.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(y.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx = (sx & 31).astype(np.float32) / np.float32(32)
    fy = (sy & 31).astype(np.float32) / np.float32(32)
    out = np.zeros(x.shape, np.float32)
    for dyy, dxx, wt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)), (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
        xx, yy = ix + dxx, iy + dyy
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        tap = np.where(ok, img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0))
        out = out + tap.astype(np.float32) * wt.astype(np.float32)
    return out


def check_geometric_consistency(depth_ref, K_ref, E_ref, depth_src, K_src, E_src, dist_base=1 / 4, rel_diff_base=1 / 1300):
    """Reference filter/dypcd.py:98-159 (reproject_with_depth + check_geometric_consistency) in numpy, same
    promotions (float64 projections, float32 where the reference casts).  Returns (masks[9], depth_reprojected)."""
    import numpy as np
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    xr, yr = x_ref.reshape([-1]), y_ref.reshape([-1])
    xyz_ref = np.matmul(np.linalg.inv(K_ref), np.vstack((xr, yr, np.ones_like(xr))) * depth_ref.reshape([-1]))
    xyz_src = np.matmul(np.matmul(E_src, np.linalg.inv(E_ref)), np.vstack((xyz_ref, np.ones_like(xr))))[:3]
    k_xyz_src = np.matmul(K_src, xyz_src)
    xy_src = k_xyz_src[:2] / k_xyz_src[2:3]
    x_src = xy_src[0].reshape([height, width]).astype(np.float32)
    y_src = xy_src[1].reshape([height, width]).astype(np.float32)
    sampled = remap_bilinear(depth_src, x_src, y_src)
    xyz_src = np.matmul(np.linalg.inv(K_src), np.vstack((xy_src, np.ones_like(xr))) * sampled.reshape([-1]))
    xyz_rep = np.matmul(np.matmul(E_ref, np.linalg.inv(E_src)), np.vstack((xyz_src, np.ones_like(xr))))[:3]
    depth_rep = xyz_rep[2].reshape([height, width]).astype(np.float32)
    k_xyz_rep = np.matmul(K_ref, xyz_rep)
    k_xyz_rep[2:3][k_xyz_rep[2:3] == 0] += 0.00001
    xy_rep = k_xyz_rep[:2] / k_xyz_rep[2:3]
    x_rep = xy_rep[0].reshape([height, width]).astype(np.float32)
    y_rep = xy_rep[1].reshape([height, width]).astype(np.float32)
    dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
    rel = np.abs(depth_rep - depth_ref) / depth_ref
    masks = [np.logical_and(dist < i * dist_base, rel < i * rel_diff_base) for i in range(2, 11)]
    depth_rep[~masks[-1]] = 0
    return masks, depth_rep


def filter_reference_view(depth_ref, confs, K_ref, E_ref, depth_srcs, K_srcs, E_srcs, conf_thr=(0.1, 0.15, 0.9),
                          dist_base=1 / 4, rel_diff_base=1 / 1300):
    """Per-reference-view part of filter_depth (reference filter/dypcd.py:205-257)."""
    import numpy as np
    photo = np.logical_and(np.logical_and(confs[2] > conf_thr[2], confs[1] > conf_thr[1]), confs[0] > conf_thr[0])
    dy_range = len(depth_srcs) + 1
    geo_sum, sums, reps = 0, [0] * (dy_range - 2), []
    for d, K, E in zip(depth_srcs, K_srcs, E_srcs):
        masks, rep = check_geometric_consistency(depth_ref, K_ref, E_ref, d, K, E, dist_base, rel_diff_base)
        geo_sum = geo_sum + masks[-1].astype(np.int32)
        for i in range(2, dy_range):
            sums[i - 2] = sums[i - 2] + masks[i - 2].astype(np.int32)
        reps.append(rep)
    avg = (sum(reps) + depth_ref) / (geo_sum + 1)
    geo = geo_sum >= dy_range
    for i in range(2, dy_range):
        geo = np.logical_or(geo, sums[i - 2] >= i)
    return {"depth_est_averaged": avg, "photo_mask": photo, "geo_mask": geo, "final_mask": np.logical_and(photo, geo)}


# --------------------------------------------------------------------------
# cross-view photometric loss (training-side neighbour of the path)
# --------------------------------------------------------------------------
def inverse_warping(img: torch.Tensor, left_cam: torch.Tensor, right_cam: torch.Tensor, depth: torch.Tensor):
    """img [B,H,W,C] of the source ("right") view warped into the reference ("left") view with the reference
    depth [B,1,H,W]; cams [B,2,4,4].  Returns (warped [B,H,W,C], mask [B,H,W,1]).

    Restates reference models/homography.py:7-201 without its ``.cuda()`` calls, quirks kept: the source pixel is
    projected with the REFERENCE intrinsics (:53-57), ``z + 1e-10`` in the perspective divide (:99-100), the
    bilinear weights use the CLAMPED x1 / y1 (:156-159, :188-191), the mask tests ``y0 <= max_y`` instead of
    ``y1`` (:153)."""
    R_l, R_r = left_cam[:, 0, :3, :3], right_cam[:, 0, :3, :3]
    t_l, t_r = left_cam[:, 0, :3, 3:4], right_cam[:, 0, :3, 3:4]
    K_l = left_cam[:, 1, :3, :3]
    K_l_inv = torch.inverse(K_l)
    R_rel = torch.matmul(R_r, R_l.permute(0, 2, 1))
    t_rel = t_r - torch.matmul(R_rel, t_l)
    b, h, w, c = img.shape
    filler = torch.tensor([0.0, 0.0, 0.0, 1.0]).reshape(1, 1, 4).repeat(b, 1, 1)
    transform = torch.cat([torch.cat([R_rel, t_rel], dim=2).float(), filler], dim=1)
    intr = torch.cat([torch.cat([K_l.float(), torch.zeros(b, 3, 1)], dim=2), filler], dim=1)
    proj = torch.matmul(intr, transform)
    x_t = torch.matmul(torch.ones(h, 1), torch.linspace(-1.0, 1.0, w).unsqueeze(1).permute(1, 0))
    y_t = torch.matmul(torch.linspace(-1.0, 1.0, h).unsqueeze(1), torch.ones(1, w))
    x_t = (x_t + 1.0) * 0.5 * (w - 1)
    y_t = (y_t + 1.0) * 0.5 * (h - 1)
    grid = torch.cat([x_t.reshape(1, -1), y_t.reshape(1, -1), torch.ones(1, h * w)], dim=0).unsqueeze(0).repeat(b, 1, 1)
    cam = torch.matmul(K_l_inv.float(), grid.float()) * depth.reshape(b, 1, h * w).float()
    pc = torch.matmul(proj, torch.cat([cam, torch.ones(b, 1, h * w)], dim=1))
    px = pc[:, 0:1] / (pc[:, 2:3] + 1e-10)
    py = pc[:, 1:2] / (pc[:, 2:3] + 1e-10)
    px = px.reshape(b, h, w, 1) / (w - 1) * 2.0 - 1.0
    py = py.reshape(b, h, w, 1) / (h - 1) * 2.0 - 1.0
    x = (px.reshape(-1).float() + 1.0) * (w - 1.0) / 2.0
    y = (py.reshape(-1).float() + 1.0) * (h - 1.0) / 2.0
    x0 = torch.floor(x).int()
    x1 = x0 + 1
    y0 = torch.floor(y).int()
    y1 = y0 + 1
    mask = ((x0 >= 0) & (x1 <= w - 1) & (y0 >= 0) & (y0 <= h - 1)).float()
    x0, x1 = torch.clamp(x0, 0, w - 1), torch.clamp(x1, 0, w - 1)
    y0, y1 = torch.clamp(y0, 0, h - 1), torch.clamp(y1, 0, h - 1)
    base = (torch.arange(b) * (w * h)).reshape(-1, 1).repeat(1, h * w).reshape(-1).long()
    flat = img.reshape(-1, c).float()
    pa, pb = flat[base + y0.long() * w + x0.long()], flat[base + y1.long() * w + x0.long()]
    pcc, pd = flat[base + y0.long() * w + x1.long()], flat[base + y1.long() * w + x1.long()]
    ux, uy = x1.float() - x, y1.float() - y
    wa, wb, wc, wd = (ux * uy).unsqueeze(1), (ux * (1.0 - uy)).unsqueeze(1), ((1.0 - ux) * uy).unsqueeze(1), ((1.0 - ux) * (1.0 - uy)).unsqueeze(1)
    out = wa * pa + wb * pb + wc * pcc + wd * pd
    return out.reshape(b, h, w, c), mask.reshape(b, h, w, 1)


def cross_view_loss(inputs, imgs: torch.Tensor, sample_cams, depth_gt_ms, depth_loss_weights) -> torch.Tensor:
    """Reference models/module.py:624-691: per stage and source view a SCALAR smooth-L1 between the source image
    warped with the estimated and with the ground-truth depth (:618-620), broadcast over the pixels where both
    warps are valid, + 1e4 elsewhere; per pixel the two smallest are summed (top-k, :672-680), averaged over the
    pixels and weighted per stage."""
    num_views = imgs.shape[1]
    total = torch.zeros(())
    for key in [k for k in inputs.keys() if "stage" in k]:
        depth_est = inputs[key]["depth"].unsqueeze(1)
        depth_gt = depth_gt_ms[key].unsqueeze(1)
        scale = depth_est.shape[-1] / imgs.shape[-1]
        ref_cam = sample_cams[key][:, 0]
        losses = []
        for v in range(1, num_views):
            view_img = F.interpolate(imgs[:, v], scale_factor=scale, mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
            view_cam = sample_cams[key][:, v]
            w_est, m_est = inverse_warping(view_img, ref_cam, view_cam, depth_est)
            w_gt, m_gt = inverse_warping(view_img, ref_cam, view_cam, depth_gt)
            mask = m_est * m_gt
            l = F.smooth_l1_loss(w_est * mask, w_gt * mask, reduction="mean")
            losses.append(l + 1e4 * (1 - mask))
        vol = torch.stack(losses).permute(1, 2, 3, 4, 0)
        top_vals, _ = torch.topk(torch.neg(vol), k=2, sorted=False)
        top_vals = torch.neg(top_vals)
        top_vals = top_vals * (top_vals < 1e4).float()
        stage_idx = int(key.replace("stage", "")) - 1
        total = total + torch.mean(torch.sum(top_vals, dim=-1)) * depth_loss_weights[stage_idx]
    return total


# --------------------------------------------------------------------------
# whole stage
# --------------------------------------------------------------------------
def depthnet_forward(stage_idx: int, features: List[torch.Tensor], proj_matrices: torch.Tensor,
                     depth_values: torch.Tensor, sd: Dict[str, torch.Tensor], mode: str = "adaptive",
                     return_volume: bool = False, training: bool = False) -> Dict[str, torch.Tensor]:
    """DepthNet.forward for one stage (reference models/cas_mvsnet.py:18-134); `training` selects batch
    statistics in every BatchNorm (model.train()).  Differentiable w.r.t. features and the tensors of `sd`
    (torch CPU autograd), which is how the tests obtain reference gradients."""
    assert depth_values.dim() == 4
    vol = aggregate(features, proj_matrices, depth_values, mode, sd, stage_idx, training)
    logits = cost_reg_net(vol, sd, stage_idx, training).squeeze(1)
    out = regress_head(logits, depth_values)
    if return_volume:
        out["volume"] = vol
        out["logits"] = logits
    return out
