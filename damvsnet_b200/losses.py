"""Cross-view photometric loss (the training-side neighbour of the hot path).

``cross_view_loss`` keeps the reference's signature and value (models/module.py:624-691, with ``inverse_warping`` of
models/homography.py:7-201 inside); per stage three native kernels replace the 2 x (N-1) warps of ~30 tensor ops each
and their autograd graph (csrc/cross_view.cu).  What stays in PyTorch is what is image-sized and input-only: the
resize of the source images to the stage resolution and the 4x4 camera products, formed exactly as the reference does.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from . import _lib
from .ops import _p, _stream


def _stage_cams(ref_cam: torch.Tensor, view_cams: Sequence[torch.Tensor]) -> torch.Tensor:
    """[B,2,4,4] cameras -> [B,n_src,21] fp32: inv(K_ref) and rows 0-2 of the reference-camera-to-source-pixel projection
    (models/homography.py:12-57; the source pixel is projected with the REFERENCE intrinsics, as there)."""
    b = ref_cam.shape[0]
    R_l, t_l, K_l = ref_cam[:, 0, :3, :3], ref_cam[:, 0, :3, 3:4], ref_cam[:, 1, :3, :3]
    K_l_inv = torch.inverse(K_l).float()
    filler = torch.tensor([0.0, 0.0, 0.0, 1.0], device=ref_cam.device).reshape(1, 1, 4).repeat(b, 1, 1)
    intr = torch.cat([torch.cat([K_l.float(), torch.zeros(b, 3, 1, device=ref_cam.device)], dim=2), filler], dim=1)
    out = []
    for cam in view_cams:
        R_r, t_r = cam[:, 0, :3, :3], cam[:, 0, :3, 3:4]
        R_rel = torch.matmul(R_r, R_l.permute(0, 2, 1))
        t_rel = t_r - torch.matmul(R_rel, t_l)
        transform = torch.cat([torch.cat([R_rel, t_rel], dim=2).float(), filler], dim=1)
        proj = torch.matmul(intr, transform)
        out.append(torch.cat([K_l_inv.reshape(b, 9), proj[:, :3].reshape(b, 12)], dim=1))
    return torch.stack(out, 1).contiguous()


class CrossViewStageFn(torch.autograd.Function):
    """One stage of the loss: (depth_est, depth_gt [B,h,w], cams [B,n,21], n source images [B,3,h,w]) -> scalar."""

    @staticmethod
    def forward(ctx, depth_est, depth_gt, cams, *imgs):
        b, h, w = depth_est.shape
        n = len(imgs)
        dev = depth_est.device
        de, dg = depth_est.detach().float().contiguous(), depth_gt.detach().float().contiguous()
        imgs = [i.detach().float().contiguous() for i in imgs]
        ptrs = (ctypes.c_void_p * n)(*[i.data_ptr() for i in imgs])
        bits = torch.empty((b, h, w), dtype=torch.int16, device=dev)
        sums = torch.zeros(n, dtype=torch.float64, device=dev)
        lib = _lib.load()
        with torch.cuda.device_of(depth_est):
            _lib.check(lib.damvs_cross_view_terms(_p(de), _p(dg), ptrs, _p(cams), b, n, h, w, _p(bits), _p(sums), _stream()))
            losses = (sums / float(b * h * w * 3)).float()
            counts = torch.zeros(n, dtype=torch.int64, device=dev)
            _lib.check(lib.damvs_cross_view_select(_p(bits), _p(losses), n, b * h * w, _p(counts), _stream()))
        share = counts.double() / float(b * h * w)
        ctx.save_for_backward(de, dg, cams, share, *imgs)
        return (share * losses.double()).sum().float()

    @staticmethod
    def backward(ctx, g):
        de, dg, cams, share, *imgs = ctx.saved_tensors
        b, h, w = de.shape
        n = len(imgs)
        coeff = (g.double() * share / float(b * h * w * 3)).float().contiguous()
        ptrs = (ctypes.c_void_p * n)(*[i.data_ptr() for i in imgs])
        g_depth = torch.empty_like(de)
        with torch.cuda.device_of(de):
            _lib.check(_lib.load().damvs_cross_view_bwd(_p(de), _p(dg), ptrs, _p(cams), _p(coeff), b, n, h, w, _p(g_depth), _stream()))
        return (g_depth, None, None) + (None,) * n


def cross_view_loss(inputs: Dict[str, Dict[str, torch.Tensor]], imgs: torch.Tensor, sample_cams: Dict[str, torch.Tensor],
                    depth_gt_ms: Dict[str, torch.Tensor], depth_loss_weights: Sequence[float]) -> torch.Tensor:
    """Reference signature (models/module.py:624): inputs = the CascadeMVSNet output dict (only the ``stageK`` entries'
    ``depth`` is used), imgs [B,N,3,H,W], sample_cams[stageK] [B,N,2,4,4], depth_gt_ms[stageK] [B,h,w]."""
    num_views = imgs.shape[1]
    if num_views < 3:
        raise ValueError("cross_view_loss sums the two smallest per-view losses: it needs at least two source views")
    total = torch.zeros((), dtype=torch.float32, device=imgs.device)
    for key in [k for k in inputs.keys() if "stage" in k]:
        depth_est = inputs[key]["depth"]
        scale = depth_est.shape[-1] / imgs.shape[-1]
        with torch.no_grad():
            views = [F.interpolate(imgs[:, v], scale_factor=scale, mode="bilinear", align_corners=True) for v in range(1, num_views)]
            cams = _stage_cams(sample_cams[key][:, 0], [sample_cams[key][:, v] for v in range(1, num_views)])
        stage_idx = int(key.replace("stage", "")) - 1
        total = total + CrossViewStageFn.apply(depth_est, depth_gt_ms[key], cams, *views) * depth_loss_weights[stage_idx]
    return total
