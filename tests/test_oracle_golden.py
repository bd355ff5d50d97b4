"""The oracle against outputs of the reference itself (committed fixtures).

No GPU.  These are the tests that pin oracle/damvs_oracle.py: every later
CUDA-vs-oracle comparison inherits its meaning from them.
"""
import pytest
import torch

from oracle import damvs_oracle as O
from tests import golden_io


def _conf_ok(got, want, tol=2e-3, frac=2e-3):
    """Confidence flips by O(p) where floor(sum p*k) sits on an integer:
    compare on a quantile, not the max norm."""
    bad = ((got - want).abs() > tol).float().mean().item()
    return bad <= frac


def test_homo_warping_matches_reference_outputs():
    g = golden_io.load_homo()
    out4 = O.homo_warping(g["src"], g["src_proj"], g["ref_proj"], g["dv4"])
    out2 = O.homo_warping(g["src"], g["src_proj"], g["ref_proj"], g["dv2"])
    assert out4.shape == g["out4"].shape
    # survey measured 2.4e-5 abs on O(1) features for this closed form
    assert (out4 - g["out4"]).abs().max().item() < 2e-4
    assert (out2 - g["out2"]).abs().max().item() < 2e-4
    assert (out4 - g["out4"]).abs().mean().item() < 2e-6


def test_depth_regression_matches_reference_outputs():
    g = golden_io.load_homo()
    torch.testing.assert_close(O.depth_regression(g["p"], g["dv4"][:, :, :5, :7]), g["reg4"], rtol=1e-6, atol=1e-4)
    torch.testing.assert_close(O.depth_regression(g["p"], g["dv2"]), g["reg2"], rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_depthnet_stage_matches_reference_outputs(mode, stage):
    sd, stages = golden_io.load_depthnet(mode)
    st = stages[stage]
    out = O.depthnet_forward(stage, st["features"], st["proj"], st["depth_values"], sd, mode, return_volume=True)
    vol_sub = out["volume"][:, :, ::2, ::3, ::3]
    scale = st["volume_sub"].abs().mean().item()
    assert (vol_sub - st["volume_sub"]).abs().max().item() < 2e-3 * max(scale, 1.0)
    assert (vol_sub - st["volume_sub"]).abs().mean().item() < 1e-5 * max(scale, 1.0)
    assert (out["logits"] - st["logits"]).abs().max().item() < 5e-3
    assert (out["prob_volume"] - st["prob_volume"]).abs().max().item() < 2e-3
    rel = ((out["depth"] - st["depth"]).abs() / st["depth"].abs().clamp_min(1.0))
    assert rel.max().item() < 1e-4, rel.max().item()
    assert torch.quantile(rel.flatten(), 0.99).item() < 2e-5
    assert _conf_ok(out["photometric_confidence"], st["photometric_confidence"])
    vrel = (out["variance"] - st["variance"]).abs() / st["variance"].abs().clamp_min(1e-2)
    assert torch.quantile(vrel.flatten(), 0.99).item() < 1e-3


def test_cost_reg_net_alone_matches_reference_logits():
    """CostRegNet.forward on the recorded volume is not available (only a sub-sample of the
    volume is stored), so feed the oracle's own volume and compare logits: already covered by
    the stage test; here check the transposed-conv + skip wiring on a tiny known case."""
    sd, _ = golden_io.load_depthnet("adaptive")
    x = torch.zeros(1, 8, 8, 8, 8)
    y = O.cost_reg_net(x, sd, 2)
    assert y.shape == (1, 1, 8, 8, 8)
    assert torch.isfinite(y).all()


REF = "/root/reference"


@pytest.mark.skipif(not __import__("os").path.isdir(REF), reason="reference checkout only exists in the build container")
def test_oracle_against_live_reference_depthnet():
    """When the reference is mounted, run its DepthNet live on fresh inputs (not the fixtures)."""
    import sys
    import warnings
    sys.path.insert(0, REF)
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "golden"))
    import make_golden
    from damvsnet_b200 import synthetic
    cas, _ = make_golden.load_reference()
    sd = synthetic.hot_path_state_dict(seed=5)
    feats, pm, dv = synthetic.make_stage_inputs(1, 1, 3, 64, 96, 16, seed=5)
    net = cas.DepthNet("adaptive", [32, 16, 8]).eval()
    from models.module import CostRegNet
    cr = CostRegNet(16, 8).eval()
    net.load_state_dict({k[len("DepthNet."):]: v for k, v in sd.items() if k.startswith("DepthNet.")})
    cr.load_state_dict({k[len("cost_regularization.1."):]: v for k, v in sd.items() if k.startswith("cost_regularization.1.")})
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = net(1, feats, pm, dv, 16, cr)
    got = O.depthnet_forward(1, feats, pm, dv, sd, "adaptive")
    rel = (got["depth"] - want["depth"]).abs() / want["depth"].abs()
    assert rel.max().item() < 1e-4
    assert (got["prob_volume"] - want["prob_volume"]).abs().max().item() < 2e-3
    assert _conf_ok(got["photometric_confidence"], want["photometric_confidence"])
    # the synthetic weights must give a non-degenerate head
    assert want["prob_volume"].max(1).values.mean().item() > 0.15


def test_hypothesis_sampling_matches_reference_fixture():
    from tests.golden_io import load_hypotheses
    fx = load_hypotheses()
    for name in ("stage2", "stage3"):
        H, W, scale, D = (int(v) for v in fx[name + "/meta"])
        cur = torch.nn.functional.interpolate(fx[name + "/depth"].unsqueeze(1), [H, W], mode="bilinear", align_corners=False)
        ev = torch.nn.functional.interpolate(fx[name + "/var"].unsqueeze(1), [H, W], mode="bilinear", align_corners=False)
        torch.testing.assert_close(O.uncertainty_aware_samples(cur, ev, D)[:, :, ::3, ::3], fx[name + "/full_sub"], rtol=1e-6, atol=1e-4)
        got = O.stage_hypotheses(fx[name + "/depth"], fx[name + "/var"], D, H, W, scale)
        torch.testing.assert_close(got, fx[name + "/out"], rtol=1e-6, atol=1e-4)
    first = O.first_stage_samples(fx["stage1/depth_values"], 48)
    want = fx["stage1/out"]
    torch.testing.assert_close(first.view(2, 48, 1, 1).expand_as(want), want, rtol=1e-6, atol=1e-4)
