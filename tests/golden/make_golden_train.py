"""Generate the training-path golden fixture: gradients produced by the REFERENCE ITSELF.

Runs only in the build container (unmodified reference imported from /root/reference, nothing copied).
For one small stage (B=2, N=3 views, C=8, D=8, 16x24) it builds the reference's ``DepthNet`` and
``CostRegNet`` (reference models/cas_mvsnet.py:10-134, models/module.py:510-541), puts them in ``train()``
mode (batch-statistics BatchNorm everywhere, incl. the per-view view-weight net), runs
``DepthNet.forward`` and back-propagates a fixed linear functional of ``depth`` and ``prob_volume`` plus
the ``variance`` output.  Stored: inputs, weights, the loss weights, the forward outputs, the gradient
of every feature map and of every parameter that receives one, and the BatchNorm running buffers after
the step.  ``adaptive`` and ``variance`` aggregation, and an ``eval()``-mode (fixed statistics) variant.

Usage:  python tests/golden/make_golden_train.py   (writes tests/golden/train_grads.npz)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)


def main():
    from make_golden import load_reference
    from damvsnet_b200 import synthetic

    torch.set_num_threads(8)
    cas, ref_module = load_reference()
    B, N, C, D, H, W = 2, 3, 8, 8, 16, 24
    blob = {}
    # stage index 2 <=> full resolution intrinsics, C = 8 (synthetic.make_stage_inputs)
    feats, pm, dv = synthetic.make_stage_inputs(2, B, N, H, W, D, seed=5)
    g = torch.Generator().manual_seed(11)
    r_depth = torch.randn(B, H, W, generator=g) * 0.1
    r_prob = torch.randn(B, D, H, W, generator=g)
    r_var = torch.rand(B, H, W, generator=g) * 0.05
    blob["features"] = torch.stack(feats, 0).numpy()
    blob["proj"] = pm.numpy()
    blob["depth_values"] = dv.numpy()
    blob["r_depth"], blob["r_prob"], blob["r_var"] = r_depth.numpy(), r_prob.numpy(), r_var.numpy()

    for mode in ("adaptive", "variance"):
        for bn_mode in ("train", "eval"):
            torch.manual_seed(3)
            depthnet = cas.DepthNet(mode=mode, in_channels=[C])
            costreg = ref_module.CostRegNet(in_channels=C, base_channels=8)
            # non-trivial BatchNorm affine parameters and running buffers
            gg = torch.Generator().manual_seed(4)
            with torch.no_grad():
                for m in list(depthnet.modules()) + list(costreg.modules()):
                    if isinstance(m, torch.nn.BatchNorm3d):
                        m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=gg))
                        m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=gg))
                        m.running_mean.copy_(0.05 * torch.randn(m.bias.shape, generator=gg))
                        m.running_var.copy_(0.3 + 0.4 * torch.rand(m.bias.shape, generator=gg))
                costreg.prob.weight.mul_(4.0)
                if mode == "adaptive":
                    wn = depthnet.weight_net[0].w_net
                    wn[0].conv.weight.abs_()
                    wn[1].conv.weight.fill_(0.9)
            depthnet.train(bn_mode == "train")
            costreg.train(bn_mode == "train")
            tag = f"{mode}/{bn_mode}/"
            sd = {}
            for k, v in depthnet.state_dict().items():
                sd["DepthNet." + k] = v.clone()
            for k, v in costreg.state_dict().items():
                sd["cost_regularization.0." + k] = v.clone()
            for k, v in sd.items():      # identical for train / eval (same seed): stored once per mode
                blob[f"{mode}/w/" + k] = v.numpy()
            fs = [f.clone().requires_grad_(True) for f in feats]
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out = depthnet(0, fs, pm, dv, D, costreg)
            loss = (out["depth"] * r_depth).sum() + (out["prob_volume"] * r_prob).sum() + (out["variance"] * r_var).sum()
            loss.backward()
            blob[tag + "loss"] = np.float64(loss.item())
            for key in ("depth", "photometric_confidence", "variance", "prob_volume"):
                blob[tag + "out/" + key] = out[key].detach().numpy()
            blob[tag + "g_features"] = torch.stack([f.grad for f in fs], 0).numpy()
            n_grad = 0
            for prefix, mod in (("DepthNet.", depthnet), ("cost_regularization.0.", costreg)):
                for k, p in mod.named_parameters():
                    if p.grad is not None:
                        blob[tag + "g/" + prefix + k] = p.grad.numpy()
                        n_grad += 1
                if bn_mode == "train":
                    for k, b in mod.named_buffers():
                        blob[tag + "buf/" + prefix + k] = b.detach().clone().numpy()
            print(tag, "loss %.6f" % loss.item(), "params with grad:", n_grad,
                  "|g_feat| %.3e" % blob[tag + "g_features"].__abs__().mean())
    np.savez_compressed(os.path.join(HERE, "train_grads.npz"), **blob)
    print("done", os.path.getsize(os.path.join(HERE, "train_grads.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
