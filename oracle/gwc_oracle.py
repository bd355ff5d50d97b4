"""Oracle of the group-wise correlation cost volume -- NOT IN THE REFERENCE.

TEST INFRASTRUCTURE ONLY.  wsmtht520/DAMVSNet aggregates by variance or adaptive per-view weights
(models/cas_mvsnet.py:14, 34-39); BASELINE.json's north star additionally names a group-wise-correlation aggregation
(configs[4], "groups 4-32").  There is no reference code to follow, so this file is an OWN restatement -- parity is
pinned only as far as its building block goes: the warp is ``oracle.damvs_oracle.homo_warping``, which IS pinned to
the reference's ``homo_warping`` outputs (tests/golden/homo_warping.npz).  The definition is the one of the group-wise
correlation literature (GwcNet, and the MVS cascades that adopted it):

    cost[b, g, d, y, x] = 1/(N-1) * sum_v  1/(C/G) * sum_{c in group g} ref[b, c, y, x] * warp_v[b, c, d, y, x]
"""
from __future__ import annotations

from typing import List

import torch

from . import damvs_oracle as O


def groupwise_correlation(features: List[torch.Tensor], proj_matrices: torch.Tensor, depth_values: torch.Tensor,
                          groups: int) -> torch.Tensor:
    """features: N x [B,C,H,W]; proj_matrices [B,N,2,4,4]; depth_values [B,D,H,W] or [B,D] -> [B,G,D,H,W]."""
    ref, srcs = features[0], features[1:]
    b, c, h, w = ref.shape
    assert c % groups == 0
    projs = torch.unbind(proj_matrices, 1)
    ref_proj = O.compose_projection(projs[0])
    d = depth_values.shape[1]
    acc = torch.zeros(b, groups, d, h, w, dtype=ref.dtype)
    for src, pm in zip(srcs, projs[1:]):
        warped = O.homo_warping(src, O.compose_projection(pm), ref_proj, depth_values)            # [B,C,D,H,W]
        prod = ref.unsqueeze(2) * warped
        acc = acc + prod.view(b, groups, c // groups, d, h, w).mean(dim=2)
    return acc / len(srcs)
