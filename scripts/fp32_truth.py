#!/usr/bin/env python
"""How far is fp32 from the truth?  The reference's own DepthNet run in float64 on the GPU is the yardstick; this package's
fp32 pipeline and the reference's fp32 path (strict: TF32 off) are both measured against it, per stage, at the full
DTU-test shape on the BN-calibrated net of the full-size parity tests.

The north star's "fp32 relative depth error <= 1e-4" compares two fp32 implementations, each with its own rounding
noise; this script separates the two contributions.

Usage (GPU box; needs oracle/_ref):  python scripts/fp32_truth.py [--out gpurun_out/fp32_truth.json] [--stages 0,1,2]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stats(a, b):
    rel = ((a.double() - b.double()).abs() / b.double().abs()).flatten()
    i = int(rel.argmax())
    k = max(1, rel.numel() // 10000)
    top = torch.topk(rel, k).values
    return {"median": rel.median().item(), "p99": torch.quantile(rel[:: max(1, rel.numel() // 2 ** 23)].float(), 0.99).item(),
            "p9999": top[-1].item(), "max": rel.max().item(), "argmax": i, "n_gt_1e-4": int((rel > 1e-4).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fp32_truth.json"))
    ap.add_argument("--stages", default="0,1,2")
    ap.add_argument("--net", default="calibrated", choices=["calibrated", "synthetic"])
    args = ap.parse_args()
    import damvsnet_b200 as dm
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner, make_workload
    from oracle import ref_loader
    from tests import golden_io
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if args.net == "calibrated":
        sharp, _ = golden_io.load_depthnet("adaptive")
        sd = {k: (v / 6.0 if k.endswith("prob.weight") else v.clone()) for k, v in sharp.items()}
    else:
        sd = synthetic.hot_path_state_dict(seed=0)
    stages = make_workload(1152, 1600, 5, [48, 32, 8], seed=0, device=dev)
    runner = HotPathRunner(sd, device=dev)
    d32, c32 = ref_loader.build_hot_path(sd, "adaptive", device=dev)
    d64, c64 = ref_loader.build_hot_path(sd, "adaptive", device=dev)
    d64, c64 = d64.double(), c64.double()
    report = {"net": args.net, "stages": {}}
    for s in [int(x) for x in args.stages.split(",")]:
        f, p, d = stages[s]
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            truth = ref_loader.float64_stage_forward(d64, c64, s, f, p, d)
            ref = d32(s, list(f), p, d, d.shape[1], c32[s])
            with dm.precision("fp32"):
                ours = runner.run_stage(s, f, p, d)
        r = {"ours_vs_truth": stats(ours["depth"], truth["depth"]), "ref_gpu_vs_truth": stats(ref["depth"], truth["depth"]),
             "ours_vs_ref_gpu": stats(ours["depth"], ref["depth"]),
             "prob_max_ours_vs_truth": (ours["prob_volume"].double() - truth["prob_volume"]).abs().max().item(),
             "prob_max_ref_vs_truth": (ref["prob_volume"].double() - truth["prob_volume"]).abs().max().item()}
        i = r["ours_vs_ref_gpu"]["argmax"]
        w = d.shape[3]
        r["worst_pixel"] = {"y": i // w, "x": i % w, "truth": truth["depth"].flatten()[i].item(), "ours": ours["depth"].flatten()[i].item(),
                            "ref": ref["depth"].flatten()[i].item(),
                            "peak_prob": truth["prob_volume"][0, :, i // w, i % w].max().item()}
        report["stages"][f"stage{s + 1}"] = r
        print(f"stage{s + 1}", json.dumps(r))
        del truth, ref, ours
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
