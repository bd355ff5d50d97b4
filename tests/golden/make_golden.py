"""Generate the golden fixtures that pin ``oracle/damvs_oracle.py``.

Runs ONLY in the build container, where the unmodified reference is mounted
read-only at /root/reference.  Nothing is copied from it: the reference package
is imported from where it lies, with the one load-time patch SURVEY.md section
8c describes (the debug prints at models/cas_mvsnet.py:275-286 index pixel
[575,1018] and crash on small images, so those lines are blanked in memory).

What is recorded (all produced by the reference's own code):
  * per stage: the inputs ``DepthNet.forward`` received inside a full
    ``CascadeMVSNet.forward`` (features, projection matrices, hypotheses), rounded
    to fp16-representable values to keep the fixture small, and the outputs of
    the reference's ``DepthNet.forward`` re-run on exactly those rounded inputs
    with hot-path weights rounded the same way;
  * the aggregated cost volume entering CostRegNet (sub-sampled) and its logits;
  * stand-alone ``homo_warping`` with [B,D] and [B,D,H,W] hypotheses.

Usage:  python tests/golden/make_golden.py   (writes tests/golden/*.npz)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def load_reference():
    sys.path.insert(0, REF)
    import models.module as ref_module  # noqa: F401  (unmodified)
    path = os.path.join(REF, "models", "cas_mvsnet.py")
    lines = open(path, encoding="utf-8").read().split("\n")
    assert lines[274].strip().startswith("if stage_idx == 2:"), lines[274]
    for i in range(274, 286):
        lines[i] = ""
    mod = types.ModuleType("models.cas_mvsnet")
    mod.__package__ = "models"
    mod.__file__ = path
    sys.modules["models.cas_mvsnet"] = mod
    exec(compile("\n".join(lines), path, "exec"), mod.__dict__)
    return mod, ref_module


def fp16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).to(torch.float32)


def main():
    from damvsnet_b200 import synthetic

    torch.manual_seed(0)
    torch.set_num_threads(8)
    cas, ref_module = load_reference()
    H, W, N = 64, 96, 5
    ndepths = [48, 32, 8]
    out_dir = HERE

    first_weights = {}
    for mode in ("adaptive", "variance"):
        torch.manual_seed(0)
        model = cas.CascadeMVSNet(refine=False, ndepths=ndepths, depth_interals_ratio=[4, 2, 1],
                                  share_cr=False, cr_base_chs=[8, 8, 8], grad_method="detach", agg_mode=mode)
        imgs = synthetic.make_images(1, N, H, W, seed=0)
        projs, intr = synthetic.make_cameras(1, N, H, W, seed=0)
        dvals = synthetic.make_depth_range(1, 192)
        # BN calibration: cumulative-average running stats from train-mode passes
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = None
        model.train()
        with torch.no_grad():
            for s in range(2):
                model(synthetic.make_images(1, N, H, W, seed=10 + s), projs, dvals, intr)
        model.eval()
        # make the random-init heads less flat so argmax-like quantities are exercised
        with torch.no_grad():
            for cr in model.cost_regularization:
                cr.prob.weight.mul_(6.0)
        # round hot-path weights to fp16-representable values
        hot = {}
        with torch.no_grad():
            for k, v in model.state_dict().items():
                if k.startswith("cost_regularization.") or k.startswith("DepthNet."):
                    if v.dtype.is_floating_point:
                        v.copy_(fp16_round(v))
                    hot[k] = v.clone()

        captured = []
        orig_forward = model.DepthNet.forward

        def spy(stage_idx, features, proj_matrices, depth_values, num_depth, cost_regularization, prob_volume_init=None):
            captured.append((stage_idx, [f.detach().clone() for f in features], proj_matrices.detach().clone(),
                             depth_values.detach().clone()))
            return orig_forward(stage_idx, features, proj_matrices, depth_values, num_depth, cost_regularization,
                                prob_volume_init)

        model.DepthNet.forward = spy
        with torch.no_grad():
            full = model(imgs, projs, dvals, intr)
        model.DepthNet.forward = orig_forward
        assert len(captured) == 3

        blob = {}
        for k, v in hot.items():
            arr = v.numpy().astype(np.float16) if v.dtype.is_floating_point else v.numpy()
            # the variance file only stores what differs from the adaptive file (BN running stats);
            # conv weights are identical because both models are built under the same seed
            if mode == "variance" and k in first_weights and np.array_equal(first_weights[k], arr):
                continue
            blob["w/" + k] = arr
        if mode == "adaptive":
            first_weights = {k[2:]: v for k, v in blob.items()}
        for stage_idx, feats, pm, dv in captured:
            feats = [fp16_round(f) for f in feats]
            # hypotheses stay fp32: fp16 spacing at ~600 mm is 0.5 mm, too coarse
            vols = []
            cr = model.cost_regularization[stage_idx]
            hook = cr.register_forward_pre_hook(lambda mod, inp: vols.append(inp[0].detach().clone()))
            logits = []
            hook2 = cr.register_forward_hook(lambda mod, inp, out: logits.append(out.detach().clone()))
            with torch.no_grad():
                out = model.DepthNet(stage_idx, feats, pm, dv, ndepths[stage_idx], cr)
            hook.remove()
            hook2.remove()
            p = f"s{stage_idx}/"
            blob[p + "features"] = torch.stack(feats, 0).numpy().astype(np.float16)   # [N,B,C,h,w]
            blob[p + "proj"] = pm.numpy()
            blob[p + "depth_values"] = dv.numpy()
            for key in ("depth", "photometric_confidence", "variance", "prob_volume"):
                blob[p + key] = out[key].numpy()
            blob[p + "logits"] = logits[0].squeeze(1).numpy()
            blob[p + "volume_sub"] = vols[0][:, :, ::2, ::3, ::3].contiguous().numpy()
            for key in ("depth", "photometric_confidence"):
                assert torch.isfinite(out[key]).all()
            pv = out["prob_volume"]
            print(mode, "stage", stage_idx, "prob max mean %.3f" % pv.max(1).values.mean().item(),
                  "depth std %.2f" % out["depth"].std().item())
        np.savez_compressed(os.path.join(out_dir, f"depthnet_{mode}.npz"), **blob)

    # stand-alone homo_warping (reference models/module.py:297)
    g = torch.Generator().manual_seed(3)
    feats, pm, dv = synthetic.make_stage_inputs(1, 2, 3, 32, 48, 6, seed=3)
    src = fp16_round(feats[1])
    sp = pm[:, 1, 0].clone()
    sp[:, :3, :4] = torch.matmul(pm[:, 1, 1, :3, :3], pm[:, 1, 0, :3, :4])
    rp = pm[:, 0, 0].clone()
    rp[:, :3, :4] = torch.matmul(pm[:, 0, 1, :3, :3], pm[:, 0, 0, :3, :4])
    dv2 = dv[:, :, 0, 0].contiguous()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w4 = ref_module.homo_warping(src, sp, rp, dv)
        w2 = ref_module.homo_warping(src, sp, rp, dv2)
    p = torch.softmax(torch.randn(2, 6, 5, 7, generator=g), 1)
    np.savez_compressed(os.path.join(out_dir, "homo_warping.npz"),
                        src=src.numpy().astype(np.float16), src_proj=sp.numpy(), ref_proj=rp.numpy(),
                        dv4=dv.numpy(), dv2=dv2.numpy(), out4=w4.numpy(), out2=w2.numpy(),
                        p=p.numpy(), reg4=ref_module.depth_regression(p, dv[:, :, :5, :7]).numpy(),
                        reg2=ref_module.depth_regression(p, dv2).numpy())
    print("done")


if __name__ == "__main__":
    main()
