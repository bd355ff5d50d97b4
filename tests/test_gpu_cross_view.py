"""GPU parity of the fused cross-view photometric loss (value and depth gradients) against the reference-generated
fixture and the oracle on a second seeded scene."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import damvs_oracle as O  # noqa: E402
from tests.golden_io import load_cross_view_loss  # noqa: E402


def dev():
    return torch.device("cuda:0")


def native(fx_imgs, cams, est, gts, w):
    from damvsnet_b200 import losses
    inputs = {k: {"depth": v.to(dev()).requires_grad_(True)} for k, v in est.items()}
    loss = losses.cross_view_loss(inputs, fx_imgs.to(dev()), {k: v.to(dev()) for k, v in cams.items()},
                                  {k: v.to(dev()) for k, v in gts.items()}, w)
    loss.backward()
    return loss.item(), {k: v["depth"].grad.cpu() for k, v in inputs.items()}


def test_matches_reference_fixture():
    fx = load_cross_view_loss()
    loss, grads = native(fx["imgs"], fx["cams"], fx["depth_est"], fx["depth_gt"], fx["dlossw"])
    assert abs(loss - fx["loss"]) <= 2e-4 * abs(fx["loss"])
    for k, want in fx["grad"].items():
        assert ((grads[k] - want).norm() / want.norm()).item() < 2e-3, k


def test_matches_oracle_second_scene():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_cvl import make_inputs
    imgs, cams, inputs, gts = make_inputs(seed=7, B=1, N=5, H=96, W=128)
    imgs = imgs + 0.2 * torch.sin(torch.arange(128.0) * 0.9).view(1, 1, 1, 1, 128)      # more texture: larger gradients
    w = [0.5, 1.0, 2.0]
    oin = {k: {"depth": v["depth"].clone().requires_grad_(True)} for k, v in inputs.items()}
    lo = O.cross_view_loss(oin, imgs, cams, gts, w)
    lo.backward()
    loss, grads = native(imgs, cams, {k: v["depth"] for k, v in inputs.items()}, gts, w)
    assert abs(loss - lo.item()) <= 2e-4 * abs(lo.item())
    for k in grads:
        want = oin[k]["depth"].grad
        assert ((grads[k] - want).norm() / want.norm()).item() < 2e-3, k
