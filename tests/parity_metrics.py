"""Error metrics shared by the GPU parity tests and scripts/ablate_precision.py.

depth_rel   |depth - ref| / |ref|                       (the north star's "relative depth error")
depth_span  |depth - ref| / (max_d hyp - min_d hyp)     (error in units of the pixel's hypothesis range: the natural
                                                          scale of a soft-argmax; stage 2/3 ranges are 24 / 6 mm wide
                                                          around ~600 mm, so depth_rel alone would hide everything there)
prob        max |prob_volume - ref|
conf        |confidence - ref|: p99 and the fraction of pixels above 2e-3 (the index floor(sum p*i) flips under 1-ulp
            changes, so a max-norm is meaningless: SURVEY.md H7)
var_rel     |variance - ref| / max(ref, 1e-2), p99
"""
from __future__ import annotations

from typing import Dict

import torch


def _q(x: torch.Tensor, q: float) -> float:
    x = x.flatten().float()
    if x.numel() > 2 ** 23:                      # torch.quantile's input limit is 16 M elements
        x = x[:: (x.numel() + 2 ** 23 - 1) // 2 ** 23]
    return torch.quantile(x, q).item()


def stage_errors(out: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], depth_values: torch.Tensor) -> Dict[str, float]:
    dev = out["depth"].device
    r = {k: v.to(dev) for k, v in ref.items() if k in ("depth", "photometric_confidence", "variance", "prob_volume")}
    dv = depth_values.to(dev)
    err = (out["depth"] - r["depth"]).abs()
    rel = err / r["depth"].abs().clamp_min(1e-6)
    span = (dv.max(1).values - dv.min(1).values).clamp_min(1e-3)
    nrm = err / span
    cerr = (out["photometric_confidence"] - r["photometric_confidence"]).abs()
    vrel = (out["variance"] - r["variance"]).abs() / r["variance"].clamp_min(1e-2)
    return {
        "depth_rel_median": rel.median().item(), "depth_rel_p99": _q(rel, 0.99), "depth_rel_max": rel.max().item(),
        "depth_span_median": nrm.median().item(), "depth_span_p99": _q(nrm, 0.99), "depth_span_max": nrm.max().item(),
        "prob_max": (out["prob_volume"] - r["prob_volume"]).abs().max().item(),
        "conf_p99": _q(cerr, 0.99), "conf_frac_gt_2e-3": (cerr > 2e-3).float().mean().item(),
        "var_rel_p99": _q(vrel, 0.99),
        "peak_prob_median": r["prob_volume"].max(1).values.median().item(),
    }
