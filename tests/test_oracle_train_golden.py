"""The oracle's training-mode restatement (batch-statistics BatchNorm, differentiated by torch CPU autograd)
against gradients produced by the reference itself (tests/golden/train_grads.npz)."""
import pytest
import torch

from oracle import damvs_oracle as O
from tests.golden_io import load_train_grads


def oracle_loss_and_grads(fx, mode, training):
    feats = [f.clone().requires_grad_(True) for f in fx["features"]]
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
          for k, v in fx["sd"].items()}
    out = O.depthnet_forward(0, feats, fx["proj"], fx["depth_values"], sd, mode, training=training)
    loss = (out["depth"] * fx["r_depth"]).sum() + (out["prob_volume"] * fx["r_prob"]).sum() + (out["variance"] * fx["r_var"]).sum()
    loss.backward()
    return out, loss, feats, sd


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("bn_mode", ["train", "eval"])
def test_oracle_gradients_match_reference(mode, bn_mode):
    fx = load_train_grads(mode, bn_mode)
    out, loss, feats, sd = oracle_loss_and_grads(fx, mode, bn_mode == "train")
    assert abs(loss.item() - fx["loss"]) <= 2e-4 * abs(fx["loss"])
    for k in ("depth", "variance", "prob_volume"):
        assert rel(out[k].detach(), fx["out"][k]) < 1e-4, k
    for f, g in zip(feats, fx["g_features"]):
        assert rel(f.grad, g) < 2e-3
    assert len(fx["grads"]) >= 31
    for k, g in fx["grads"].items():
        assert sd[k].grad is not None, k
        # w_net.1.conv.weight feeds a BatchNorm directly: its true gradient is zero up to eps effects (|g| ~ 1e-4)
        assert rel(sd[k].grad, g) < 5e-3 or float((sd[k].grad - g).abs().max()) < 2e-5, (k, rel(sd[k].grad, g))
    # parameters the reference leaves without gradient stay without one (dead conv0 of the weight net)
    for k, v in sd.items():
        if v.requires_grad and k not in fx["grads"]:
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
