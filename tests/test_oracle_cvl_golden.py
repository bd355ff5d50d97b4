"""The oracle's restatement of the cross-view photometric loss against the loss and depth gradients produced by the
reference's own cross_view_loss / inverse_warping (tests/golden/cross_view_loss.npz)."""
import torch

from oracle import damvs_oracle as O
from tests.golden_io import load_cross_view_loss


def test_cross_view_loss_and_gradients_match_reference():
    fx = load_cross_view_loss()
    inputs = {k: {"depth": v.clone().requires_grad_(True)} for k, v in fx["depth_est"].items()}
    loss = O.cross_view_loss(inputs, fx["imgs"], fx["cams"], fx["depth_gt"], fx["dlossw"])
    loss.backward()
    assert abs(loss.item() - fx["loss"]) <= 1e-5 * abs(fx["loss"])
    for k, want in fx["grad"].items():
        got = inputs[k]["depth"].grad
        assert ((got - want).norm() / want.norm()).item() < 1e-4, k
