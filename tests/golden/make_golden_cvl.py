"""Golden fixture for the cross-view photometric loss, produced by the REFERENCE's own
``cross_view_loss`` (models/module.py:624-691) and ``inverse_warping`` (models/homography.py).  The reference hard-codes
``.cuda()`` in both; the build container has no GPU, so ``torch.Tensor.cuda`` is patched to the identity for the run
(a load-time patch of the environment, not of the reference).  Stored: inputs, the loss, and its gradient with respect
to every stage's estimated depth.   Usage: python tests/golden/make_golden_cvl.py  (writes cross_view_loss.npz)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)


def make_inputs(seed=0, B=2, N=4, H=64, W=96):
    from damvsnet_b200 import synthetic
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    base = torch.stack([torch.sin(6 * xx + 3 * yy), torch.cos(5 * yy - 2 * xx), torch.sin(4 * xx * yy + 1)], 0)
    imgs = (0.5 + 0.4 * base.unsqueeze(0).unsqueeze(0) + 0.05 * torch.randn(B, N, 3, H, W, generator=g)).contiguous()
    cams, _ = synthetic.make_cameras(B, N, H, W, seed=seed + 3, max_angle=0.03, max_trans=25.0)
    inputs, gts = {}, {}
    for s, sc in enumerate((4, 2, 1)):
        h, w = H // sc, W // sc
        surf = 600 + 60 * torch.sin(3 * torch.linspace(0, 1, w)).view(1, 1, w) * torch.cos(2 * torch.linspace(0, 1, h)).view(1, h, 1)
        gt = surf.expand(B, h, w).contiguous() + torch.randn(B, h, w, generator=g)
        est = gt + 4.0 * torch.randn(B, h, w, generator=g)
        est[:, : h // 8] += 300.0          # pushes part of the image out of the source views (mask = 0)
        inputs[f"stage{s + 1}"] = {"depth": est}
        gts[f"stage{s + 1}"] = gt
    return imgs, cams, inputs, gts


def main():
    from make_golden import load_reference
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference calls .cuda() unconditionally
    cas, ref_module = load_reference()
    imgs, cams, inputs, gts = make_inputs()
    for k in inputs:
        inputs[k]["depth"].requires_grad_(True)
    w = [0.5, 1.0, 2.0]
    loss = ref_module.cross_view_loss(inputs, imgs, cams, gts, w)
    loss.backward()
    blob = {"imgs": imgs.numpy(), "loss": np.float64(loss.item()), "dlossw": np.array(w)}
    for k in inputs:
        blob[k + "/cams"] = cams[k].numpy()
        blob[k + "/depth_est"] = inputs[k]["depth"].detach().numpy()
        blob[k + "/depth_gt"] = gts[k].numpy()
        blob[k + "/grad"] = inputs[k]["depth"].grad.numpy()
    np.savez_compressed(os.path.join(HERE, "cross_view_loss.npz"), **blob)
    print("loss", loss.item(), "|grad|", [float(inputs[k]["depth"].grad.abs().sum()) for k in inputs],
          os.path.getsize(os.path.join(HERE, "cross_view_loss.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
