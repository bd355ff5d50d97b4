"""Seeded synthetic inputs for the cost-volume hot path.

Everything here is plain torch on the CPU; callers move tensors where they need
them.  No dataset is available offline, so cameras follow the DTU conventions
the reference's loaders produce (datasets/general_eval.py:157-180,
datasets/dtu_yao.py:65-66,202-203): ``proj[:, v, 0]`` is a 4x4 extrinsic,
``proj[:, v, 1, :3, :3]`` the intrinsic at that stage's resolution, and the
stage-2/3 intrinsics are the stage-1 ones with rows 0-1 multiplied by 2 and 4.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

STAGE_SCALES = (4, 2, 1)          # models/cas_mvsnet.py:154-164
STAGE_CHANNELS = (32, 16, 8)      # models/module.py:383,408-409 (FPN, base 8)
DTU_DEPTH_MIN = 425.0
DTU_DEPTH_INTERVAL = 2.65         # 2.5 * 1.06


def _rot(ax: float, ay: float, az: float) -> np.ndarray:
    cx, sx = math.cos(ax), math.sin(ax)
    cy, sy = math.cos(ay), math.sin(ay)
    cz, sz = math.cos(az), math.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return rz @ ry @ rx


def make_cameras(batch: int, nviews: int, height: int, width: int, seed: int = 0,
                 max_angle: float = 0.08, max_trans: float = 60.0) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Return (proj_matrices, intrinsics_matrices) dicts keyed stage1..stage3.

    proj_matrices[stage] is [B, N, 2, 4, 4] fp32, intrinsics_matrices[stage] is
    [B, 3, 3] fp32 (the reference view's K), as the reference's datasets emit.
    """
    rs = np.random.RandomState(seed)
    f = 2892.33 * width / 1600.0
    k1 = np.array([[f / 4.0, 0, width / 8.0 - 0.5],
                   [0, f / 4.0, height / 8.0 - 0.5],
                   [0, 0, 1]], dtype=np.float64)
    proj = np.zeros((batch, nviews, 2, 4, 4), dtype=np.float32)
    for b in range(batch):
        for v in range(nviews):
            ext = np.eye(4, dtype=np.float64)
            if v > 0:
                ang = rs.uniform(-max_angle, max_angle, size=3)
                ext[:3, :3] = _rot(*ang)
                ext[:3, 3] = rs.uniform(-max_trans, max_trans, size=3)
            proj[b, v, 0] = ext.astype(np.float32)
            proj[b, v, 1, :3, :3] = k1.astype(np.float32)
    projs, intr = {}, {}
    for i, mul in enumerate((1.0, 2.0, 4.0)):
        p = proj.copy()
        p[:, :, 1, :2, :] = proj[:, :, 1, :2, :] * mul
        projs[f"stage{i + 1}"] = torch.from_numpy(p)
        intr[f"stage{i + 1}"] = torch.from_numpy(p[:, 0, 1, :3, :3].copy())
    return projs, intr


def make_depth_range(batch: int, numdepth: int = 192) -> torch.Tensor:
    """[B, numdepth] plane sweep range, DTU defaults (datasets/dtu_yao.py:202-203)."""
    d = DTU_DEPTH_MIN + DTU_DEPTH_INTERVAL * torch.arange(numdepth, dtype=torch.float32)
    return d.unsqueeze(0).repeat(batch, 1).contiguous()


def make_images(batch: int, nviews: int, height: int, width: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, nviews, 3, height, width, generator=g, dtype=torch.float32)


def smooth_features(batch: int, channels: int, h: int, w: int, g: torch.Generator,
                    shared: torch.Tensor | None = None, mix: float = 0.7) -> torch.Tensor:
    """O(1) features with some spatial structure; ``shared`` correlates views so
    the cost volume is not pure noise."""
    coarse = torch.randn(batch, channels, max(h // 4, 1), max(w // 4, 1), generator=g)
    up = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)
    x = 0.6 * up + 0.4 * torch.randn(batch, channels, h, w, generator=g)
    if shared is not None:
        x = mix * shared + (1.0 - mix) * x
    return x.contiguous()


def make_stage_inputs(stage_idx: int, batch: int, nviews: int, height: int, width: int,
                      ndepth: int, seed: int = 0, per_pixel_jitter: bool = True,
                      channels: int | None = None):
    """Direct inputs of DepthNet.forward for one stage (models/cas_mvsnet.py:18).

    Returns (features: list of N [B,C,h,w], proj_matrices [B,N,2,4,4],
    depth_values [B,D,h,w]).  Hypotheses are per pixel, as the live caller
    always passes them (models/cas_mvsnet.py:293-296).
    """
    s = STAGE_SCALES[stage_idx]
    c = channels if channels is not None else STAGE_CHANNELS[stage_idx]
    h, w = height // s, width // s
    g = torch.Generator().manual_seed(seed * 1000 + stage_idx)
    projs, _ = make_cameras(batch, nviews, height, width, seed=seed)
    base = smooth_features(batch, c, h, w, g)
    feats: List[torch.Tensor] = [base]
    for _ in range(1, nviews):
        feats.append(smooth_features(batch, c, h, w, g, shared=base))
    if stage_idx == 0:
        lo, hi = DTU_DEPTH_MIN, DTU_DEPTH_MIN + DTU_DEPTH_INTERVAL * 191
    else:
        # later stages sweep a narrow band around a smooth surface
        half = 12.0 if stage_idx == 1 else 3.0
        lo, hi = -half, half
    lin = torch.linspace(lo, hi, ndepth, dtype=torch.float32).view(1, ndepth, 1, 1)
    if stage_idx == 0:
        dv = lin.expand(batch, ndepth, h, w).clone()
    else:
        yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
        surf = 600.0 + 120.0 * torch.sin(3.0 * xx + 0.5) * torch.cos(2.0 * yy)
        dv = surf.view(1, 1, h, w) + lin
        dv = dv.expand(batch, ndepth, h, w).clone()
    if per_pixel_jitter:
        interval = (hi - lo) / max(ndepth - 1, 1)
        dv = dv + (torch.rand(batch, 1, h, w, generator=g) - 0.5) * 0.5 * interval
    return feats, projs[f"stage{stage_idx + 1}"], dv.contiguous()


def make_prev_maps(stage_idx: int, batch: int, height: int, width: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Smooth synthetic (depth, variance) maps [B,hp,wp] of the stage BEFORE `stage_idx` (1 or 2), at that stage's
    resolution: what CascadeMVSNet.forward feeds to uncertainty_aware_samples (reference models/cas_mvsnet.py:236-274).
    The surface and the half-ranges (12 mm / 3 mm) are those of make_stage_inputs, so the sampled hypotheses sweep the
    same band; `seed` shifts the phase so ranks / views differ."""
    assert stage_idx in (1, 2)
    sp = STAGE_SCALES[stage_idx - 1]
    hp, wp = height // sp, width // sp
    yy, xx = torch.meshgrid(torch.linspace(0, 1, hp), torch.linspace(0, 1, wp), indexing="ij")
    ph = 0.37 * seed
    surf = 600.0 + 120.0 * torch.sin(3.0 * xx + 0.5 + ph) * torch.cos(2.0 * yy + 0.5 * ph)
    half = 12.0 if stage_idx == 1 else 3.0
    var = half * (1.0 + 0.1 * torch.sin(5.0 * xx + ph) * torch.cos(4.0 * yy))
    return (surf.unsqueeze(0).repeat(batch, 1, 1).contiguous(), var.unsqueeze(0).repeat(batch, 1, 1).contiguous())


def hot_path_state_dict(in_channels=STAGE_CHANNELS, base_channels=(8, 8, 8), seed: int = 0,
                        mode: str = "adaptive") -> Dict[str, torch.Tensor]:
    """Random, *non-degenerate* parameters for the hot path under the reference's
    state_dict keys: cost_regularization.{s}.* and
    DepthNet.weight_net.{s}.*.  BatchNorm running stats are drawn so that eval
    mode produces peaked probability volumes (SURVEY.md section 0.5 explains
    why default init + eval() is useless for parity).
    """
    g = torch.Generator().manual_seed(seed + 7)
    sd: Dict[str, torch.Tensor] = {}

    def conv_w(cout, cin, k, transposed=False):
        fan_in = cin * k ** 3
        bound = 1.0 / math.sqrt(fan_in)
        shape = (cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound * math.sqrt(3.0)

    def bn(prefix, c, mean_scale=0.05, var_lo=0.15, var_hi=0.45):
        sd[prefix + ".weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[prefix + ".running_mean"] = mean_scale * torch.randn(c, generator=g)
        sd[prefix + ".running_var"] = var_lo + (var_hi - var_lo) * torch.rand(c, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    for s, (cin, b) in enumerate(zip(in_channels, base_channels)):
        p = f"cost_regularization.{s}."
        plan = [("conv0", cin, b, False), ("conv1", b, 2 * b, False), ("conv2", 2 * b, 2 * b, False),
                ("conv3", 2 * b, 4 * b, False), ("conv4", 4 * b, 4 * b, False),
                ("conv5", 4 * b, 8 * b, False), ("conv6", 8 * b, 8 * b, False),
                ("conv7", 8 * b, 4 * b, True), ("conv9", 4 * b, 2 * b, True), ("conv11", 2 * b, b, True)]
        for name, ci, co, tr in plan:
            sd[p + name + ".conv.weight"] = conv_w(co, ci, 3, tr)
            bn(p + name + ".bn", co)
        sd[p + "prob.weight"] = conv_w(1, b, 3) * 4.0
        if mode == "adaptive":
            q = f"DepthNet.weight_net.{s}."
            sd[q + "conv0.conv.weight"] = conv_w(1, cin, 1)
            bn(q + "conv0.bn", 1)
            sd[q + "w_net.0.conv.weight"] = conv_w(1, cin, 1).abs()
            bn(q + "w_net.0.bn", 1, mean_scale=0.3, var_lo=0.2, var_hi=0.6)
            sd[q + "w_net.1.conv.weight"] = torch.rand(1, 1, 1, 1, 1, generator=g) + 0.5
            bn(q + "w_net.1.bn", 1, mean_scale=0.3, var_lo=0.2, var_hi=0.6)
    return sd
