#!/usr/bin/env python
"""Where does the reduced-precision pipeline's error come from?  ("what's weak" 1a.)

Teacher-forced per stage at the FULL DTU-test shape (1152x1600, N=5, D=48/32/8), every row against this package's own
fp32 pipeline on the same inputs (that pipeline is pinned to the reference at <= 1e-4 relative depth error by the parity
tests, at fixture size and at full size), for three nets:

  calibrated      the BN-calibrated random-init net of tests/golden/depthnet_adaptive.npz with its x6 head sharpening undone
                  (SURVEY.md H7's setting: "un-sharpened, BN-calibrated")
  sharpened       the same fixture as committed (prob.weight x 6: peaked, multi-modal probability volumes)
  synthetic       damvsnet_b200.synthetic.hot_path_state_dict (the net bench.py runs)

Rows (what is rounded):
  A  vol_bf16            fp32 features, cost volume rounded to bf16, fp32 direct convolutions
  B  feat_fp16           fp16 features (fp32 blend/accumulate), fp32 cost volume, fp32 direct convolutions
  C  conv_bf16           fp32 features, bf16 cost volume, tcgen05 convolutions with bf16 weights/activations  (north star's
                         "bf16 on the tensor-core convs")
  D  bf16 pipeline       fp16 features, bf16 cost volume, tcgen05 convolutions  (precision "bf16": round 1's benchmarked mode)
  E  vol_fp16            fp32 features, cost volume rounded to fp16, fp32 direct convolutions
  F  conv_fp16           fp32 features, fp16 cost volume, tcgen05 convolutions with fp16 weights/activations
  G  fp16 pipeline       fp16 features, fp16 cost volume, tcgen05 convolutions  (precision "fp16": bench.py's default)

Usage (GPU box):  python scripts/ablate_precision.py [--out gpurun_out/ablation.json] [--small]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nets():
    from damvsnet_b200 import synthetic
    from tests import golden_io
    sharp, _ = golden_io.load_depthnet("adaptive")
    calib = {k: (v / 6.0 if k.endswith("prob.weight") else v.clone()) for k, v in sharp.items()}
    return {"calibrated": calib, "sharpened": sharp, "synthetic": synthetic.hot_path_state_dict(seed=0)}


@torch.no_grad()
def run_stage(dm, runner, s, feats, proj, dv, feat_half: bool, vol_dtype, conv_half):
    from damvsnet_b200 import ops
    net, cr = runner.depthnet, runner.cost_regularization[s]
    rot_trans = net.stage_rot_trans(proj)
    nhwc = ops.features_to_nhwc_half_multi(feats) if feat_half else [ops.features_to_nhwc(f) for f in feats]
    wn = net.weight_net[s].folded()
    vol = ops.warp_aggregate(nhwc[0], nhwc[1:], rot_trans, dv, wn, "adaptive", vol_dtype)
    if conv_half:       # "bf16" | "fp16": tcgen05 convolutions on volumes / weights of that type
        hd = torch.bfloat16 if conv_half == "bf16" else torch.float16
        with dm.precision(conv_half):
            logits = cr.forward_g8(vol if vol.dtype == hd else ops.G8Volume(vol.data.to(hd)))
    else:
        with dm.precision("fp32"):
            logits = cr.forward_g8(vol if vol.dtype == torch.float32 else ops.G8Volume(vol.data.float()))
    prob, depth, conf, var = ops.softmax_regress(logits, dv)
    return {"depth": depth, "photometric_confidence": conf, "variance": var, "prob_volume": prob}


ROWS = (("A vol_bf16", False, torch.bfloat16, None),
        ("B feat_fp16", True, torch.float32, None),
        ("C conv_bf16", False, torch.bfloat16, "bf16"),
        ("D bf16 pipeline", True, torch.bfloat16, "bf16"),
        ("E vol_fp16", False, torch.float16, None),
        ("F conv_fp16", False, torch.float16, "fp16"),
        ("G fp16 pipeline", True, torch.float16, "fp16"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ablation.json"))
    ap.add_argument("--small", action="store_true", help="288x400 instead of 1152x1600 (quick check)")
    args = ap.parse_args()
    import damvsnet_b200 as dm
    from damvsnet_b200.runner import HotPathRunner, make_workload
    from tests.parity_metrics import stage_errors
    dev = torch.device("cuda:0")
    H, W = (288, 400) if args.small else (1152, 1600)
    stages = make_workload(H, W, 5, [48, 32, 8], seed=0, device=dev)
    report = {"shape": [H, W], "nviews": 5, "ndepths": [48, 32, 8], "baseline": "this package, precision fp32", "nets": {}}
    for name, sd in nets().items():
        runner = HotPathRunner(sd, device=dev)
        rep = {}
        for s, (feats, proj, dv) in enumerate(stages):
            base = run_stage(dm, runner, s, feats, proj, dv, False, torch.float32, False)
            for tag, fh, vd, cb in ROWS:
                out = run_stage(dm, runner, s, feats, proj, dv, fh, vd, cb)
                rep.setdefault(tag, {})[f"stage{s + 1}"] = stage_errors(out, base, dv)
                del out
            del base
            torch.cuda.empty_cache()
        report["nets"][name] = rep
        del runner
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)
    # markdown table
    keys = ("depth_rel_median", "depth_rel_p99", "depth_rel_max", "depth_span_median", "depth_span_p99", "prob_max", "conf_p99")
    print("| net | row | stage | " + " | ".join(keys) + " | peak prob (median) |")
    print("|---|---|---|" + "---|" * (len(keys) + 1))
    for name, rep in report["nets"].items():
        for tag, st in rep.items():
            for sname, e in st.items():
                print(f"| {name} | {tag} | {sname} | " + " | ".join(f"{e[k]:.2e}" for k in keys) + f" | {e['peak_prob_median']:.3f} |")


if __name__ == "__main__":
    main()
