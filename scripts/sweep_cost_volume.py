#!/usr/bin/env python
"""Cost-volume microbenchmark sweep (BASELINE.json configs[4]): D in {48, 96, 192}, C in {8, 16, 32}, resolutions up
to 2048x2560, N = 5 views; device time and roofline fraction of each hot-path kernel in isolation (CUDA events around
graph-captured launches, inputs resident), precision fp16.  The "groups" axis: the reference itself has no group-wise
correlation -- its variance / adaptive aggregations are swept over the channel width C, and the
group-wise correlation variant this package adds (csrc/warp_gwc.cu) over G in {4, 8, 16, 32} (G <= C).

    python scripts/sweep_cost_volume.py [--out profiles/r02_sweep.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sweep.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import damvsnet_b200 as dm
    dm.set_precision("fp16")   # measures the reduced-precision pipeline (the package default is fp32)
    from damvsnet_b200 import ops, synthetic
    from damvsnet_b200.runner import HotPathRunner
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    sd = synthetic.hot_path_state_dict(seed=0)
    runner = HotPathRunner(sd, device=dev)
    N = 5
    rows = []
    shapes = [(32, 0, 288, 400), (16, 1, 576, 800), (8, 2, 1152, 1600), (8, 2, 2048, 2560), (16, 1, 1024, 1280), (32, 0, 512, 640)]
    depths = [48, 96, 192]
    if args.quick:
        shapes, depths = shapes[:3], [48]
    for C, stage, h, w in shapes:
        for D in depths:
            vox = D * h * w
            if vox * C * 2 > 40e9 or vox > 1.1e9:
                continue
            g = torch.Generator().manual_seed(0)
            feats = [torch.randn(1, C, h, w, generator=g).to(dev) for _ in range(N)]
            projs, _ = synthetic.make_cameras(1, N, h * 4 // (4 >> min(stage, 2)) if False else h * synthetic.STAGE_SCALES[stage], w * synthetic.STAGE_SCALES[stage], seed=0)
            pm = projs[f"stage{stage + 1}"].to(dev)
            dv = (425 + (506.0 / max(D - 1, 1)) * torch.arange(D, dtype=torch.float32).view(1, D, 1, 1) + torch.rand(1, 1, h, w, generator=g)).expand(1, D, h, w).contiguous().to(dev)
            net, cr = runner.depthnet, runner.cost_regularization[stage]
            rt = net.stage_rot_trans(pm)
            nhwc = [ops.features_to_nhwc_half(f) for f in feats]
            wnet = net.weight_net[stage].folded()
            vol = ops.warp_aggregate(nhwc[0], nhwc[1:], rt, dv, wnet, "adaptive", torch.float16)
            t_warp = timed(lambda: ops.warp_aggregate(nhwc[0], nhwc[1:], rt, dv, wnet, "adaptive", torch.float16))
            t_var = timed(lambda: ops.warp_aggregate(nhwc[0], nhwc[1:], rt, dv, None, "variance", torch.float16))
            b_warp = N * C * h * w * 2 + vox * 4 + vox * C * 2        # fp16 features read, fp32 hypotheses, fp16 volume written
            gwc = {}
            for G in (4, 8, 16, 32):
                if G > C:
                    continue
                t_g = timed(lambda: ops.warp_groupwise(nhwc[0], nhwc[1:], rt, dv, G, torch.float16))
                b_g = N * C * h * w * 2 + vox * 4 + vox * max(G, 8) * 2
                gwc[f"G{G}"] = {"ms": t_g, "GBps": b_g / t_g / 1e6, "hbm_frac": b_g / t_g / 1e6 / peaks["hbm_gbs"]}
            logits = cr.forward_g8(vol)
            t_reg = timed(lambda: cr.forward_g8(vol))
            chans = [(C, 8, 1.0), (8, 16, 1 / 8), (16, 16, 1 / 8), (16, 32, 1 / 64), (32, 32, 1 / 64), (32, 64, 1 / 512), (64, 64, 1 / 512),
                     (64, 32, 1 / 512), (32, 16, 1 / 64), (16, 8, 1 / 8), (8, 1, 1.0)]
            flops = sum(2 * 27 * ci * co * vox * f for ci, co, f in chans)
            t_head = timed(lambda: ops.softmax_regress(logits, dv))
            b_head = 2 * vox * 4 + vox * 4 + 3 * h * w * 4
            rows.append({"C": C, "D": D, "h": h, "w": w, "voxels": vox,
                         "warp_agg_adaptive_ms": t_warp, "warp_agg_adaptive_GBps": b_warp / t_warp / 1e6, "warp_agg_adaptive_hbm_frac": b_warp / t_warp / 1e6 / peaks["hbm_gbs"],
                         "warp_agg_variance_ms": t_var, "warp_agg_variance_GBps": b_warp / t_var / 1e6, "warp_gwc": gwc,
                         "costregnet_ms": t_reg, "costregnet_TFLOPs": flops / t_reg / 1e9, "costregnet_tensor_frac": flops / t_reg / 1e9 / peaks["bf16_tflops_sustained"],
                         "head_ms": t_head, "head_GBps": b_head / t_head / 1e6, "head_hbm_frac": b_head / t_head / 1e6 / peaks["hbm_gbs"]})
            r = rows[-1]
            print(f"C={C:2d} D={D:3d} {h}x{w}: warp {t_warp:7.3f} ms ({r['warp_agg_adaptive_GBps']:6.0f} GB/s)  variance {t_var:7.3f}  gwc {' '.join(f'{k}:{v["ms"]:.3f}' for k, v in gwc.items())}  costreg {t_reg:7.3f} ms ({r['costregnet_TFLOPs']:5.0f} TF/s)  head {t_head:6.3f} ms ({r['head_GBps']:5.0f} GB/s)", flush=True)
            del vol, logits, feats, nhwc, dv
            torch.cuda.empty_cache()
    json.dump({"peaks": peaks, "n_views": N, "precision": "fp16", "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
