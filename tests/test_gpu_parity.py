"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
committed reference outputs.  Run on the B200 box with ``-m gpu``.

Tolerances: fp32 mode (the package default) -- relative depth error <= 1e-4 (north star).  bf16 mode (fp16 features,
bf16 cost volume, tcgen05 convolutions; opt-in) -- on these 64x96 fixtures, whose heads are sharpened x6 to make the
probability volumes multi-modal, the error is stated in units of the per-pixel hypothesis span (see
test_depthnet_bf16_within_stated_bound); the bound in relative depth (p99 <= 1e-3, max <= 5e-3) is
checked at full size on the un-sharpened net by tests/test_gpu_fullsize.py, and DESIGN.md section 5 holds the
per-component ablation.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import damvs_oracle as O  # noqa: E402
from tests import golden_io  # noqa: E402


def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def dm():
    import damvsnet_b200 as dm
    from damvsnet_b200 import _lib
    _lib.check(_lib.load().damvs_check_device(0))
    return dm


def _rel(a, b, floor=1.0):
    return ((a - b).abs() / b.abs().clamp_min(floor))


def _build_net(dm, sd, stage, mode):
    from damvsnet_b200 import synthetic
    cin = synthetic.STAGE_CHANNELS[stage]
    net = dm.DepthNet(mode, list(synthetic.STAGE_CHANNELS)).eval()
    cr = dm.CostRegNet(cin, 8).eval()
    if mode == "adaptive":
        net.load_state_dict({k[len("DepthNet."):]: v for k, v in sd.items() if k.startswith("DepthNet.")}, strict=True)
    pre = f"cost_regularization.{stage}."
    cr.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    return net.to(dev()), cr.to(dev())


# ---------------------------------------------------------------- head
@pytest.mark.parametrize("D", [8, 16, 32, 48, 64, 96])
@pytest.mark.parametrize("per_pixel", [True, False])
def test_head_matches_oracle(dm, D, per_pixel):
    g = torch.Generator().manual_seed(D)
    B, H, W = 2, 24, 40
    logits = torch.randn(B, D, H, W, generator=g) * 3
    hyp = 425 + 2.65 * torch.arange(D, dtype=torch.float32).view(1, D, 1, 1) + torch.rand(B, 1, H, W, generator=g)
    hyp = hyp.expand(B, D, H, W).contiguous()
    want = O.regress_head(logits, hyp)
    dv = hyp if per_pixel else hyp[:, :, 0, 0].contiguous()
    if not per_pixel:
        want = O.regress_head(logits, dv.view(B, D, 1, 1).expand(B, D, H, W).contiguous())
    prob, depth, conf, var = dm.ops.softmax_regress(logits.to(dev()), dv.to(dev()))
    if (want["prob_volume"] - torch.softmax(logits, 1)).abs().max() > 1e-6:
        # Seen intermittently on the GPU boxes' hosts when this runs right after the large multi-threaded CPU convolutions
        # of test_gpu_conv_tc.py: the oracle's CPU result disagreed with torch.softmax of the same logits by 3.6e-5 while
        # the CUDA result matched torch's softmax on the GPU bit for bit.  The oracle is the checker: recompute it and
        # insist that it is self-consistent before judging the kernel against it.
        want = O.regress_head(logits, hyp if per_pixel else dv.view(B, D, 1, 1).expand(B, D, H, W).contiguous())
        assert (want["prob_volume"] - torch.softmax(logits, 1)).abs().max() <= 1e-6, "CPU oracle is not self-consistent"
    assert (prob.cpu() - want["prob_volume"]).abs().max() < 1e-6
    assert _rel(depth.cpu(), want["depth"]).max() < 2e-6
    assert ((conf.cpu() - want["photometric_confidence"]).abs() > 1e-5).float().mean() < 2e-3
    assert _rel(var.cpu(), want["variance"], 1e-2).max() < 1e-3
    assert (prob.sum(1) - 1).abs().max() < 1e-5


def test_depth_regression_signature(dm):
    g = golden_io.load_homo()
    got4 = dm.depth_regression(g["p"].to(dev()), g["dv4"][:, :, :5, :7].contiguous().to(dev()))
    got2 = dm.depth_regression(g["p"].to(dev()), g["dv2"].to(dev()))
    torch.testing.assert_close(got4.cpu(), g["reg4"], rtol=1e-6, atol=1e-4)
    torch.testing.assert_close(got2.cpu(), g["reg2"], rtol=1e-6, atol=1e-4)


# ---------------------------------------------------------------- warp
def test_homo_warping_matches_reference_fixture(dm):
    g = golden_io.load_homo()
    for dv, want in ((g["dv4"], g["out4"]), (g["dv2"], g["out2"])):
        got = dm.homo_warping(g["src"].to(dev()), g["src_proj"].to(dev()), g["ref_proj"].to(dev()), dv.to(dev()))
        assert got.shape == want.shape
        assert (got.cpu() - want).abs().max() < 2e-4
        assert (got.cpu() - want).abs().mean() < 2e-6


def test_homo_warping_identity_is_identity(dm):
    """Size-independent property: with src_proj == ref_proj the effective sample coordinate is
    x*W/(W-1) - 0.5 (the reference's normalisation quirk), independent of depth."""
    B, C, H, W, D = 1, 8, 33, 47, 3
    src = torch.randn(B, C, H, W)
    P = torch.eye(4).unsqueeze(0)
    P[:, 0, 0] = 700.0
    P[:, 1, 1] = 700.0
    P[:, 0, 2] = 20.0
    P[:, 1, 2] = 15.0
    dv = torch.tensor([[400.0, 600.0, 900.0]])
    got = dm.homo_warping(src.to(dev()), P.to(dev()), P.to(dev()), dv.to(dev())).cpu()
    want = O.homo_warping(src, P, P, dv)
    assert (got - want).abs().max() < 1e-4
    assert (got[:, :, 0] - got[:, :, 2]).abs().max() < 1e-4


# ---------------------------------------------------------------- layout
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_g8_round_trip(dm, dtype):
    x = torch.randn(2, 16, 5, 9, 13)
    if dtype != torch.float32:
        x = x.to(dtype).float()
    vol = dm.G8Volume.from_ncdhw(x.to(dev()), dtype)
    assert vol.data.shape == (2, 2, 5, 9, 13, 8)
    ref = x.view(2, 2, 8, 5, 9, 13).permute(0, 1, 3, 4, 5, 2)
    assert torch.equal(vol.data.float().cpu(), ref)
    assert torch.equal(vol.to_ncdhw().cpu(), x)


def test_nhwc_repack_and_zero_copy(dm):
    x = torch.randn(2, 32, 19, 23, device=dev())
    y = dm.ops.features_to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    xc = x.contiguous(memory_format=torch.channels_last)
    yc = dm.ops.features_to_nhwc(xc)
    assert yc.data_ptr() == xc.data_ptr() and torch.equal(yc, y)


# ---------------------------------------------------------------- aggregate
@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_cost_volume_matches_reference_fixture(dm, mode, stage):
    sd, stages = golden_io.load_depthnet(mode)
    st = stages[stage]
    net, _ = _build_net(dm, sd, stage, mode)
    vol = net.cost_volume(stage, [f.to(dev()) for f in st["features"]], st["proj"].to(dev()),
                          st["depth_values"].to(dev()), out_dtype=torch.float32)
    got = vol.to_ncdhw().cpu()[:, :, ::2, ::3, ::3]
    scale = max(st["volume_sub"].abs().mean().item(), 1.0)
    assert (got - st["volume_sub"]).abs().max() < 2e-3 * scale
    assert (got - st["volume_sub"]).abs().mean() < 1e-5 * scale


@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("C", [8, 16, 32])
def test_cost_volume_matches_oracle_ragged(dm, mode, C):
    """Ragged extents (not multiples of the CTA tile), 2 batch items, [B,D] and [B,D,H,W] hypotheses."""
    from damvsnet_b200 import synthetic
    stage = {32: 0, 16: 1, 8: 2}[C]
    sd = synthetic.hot_path_state_dict(seed=11)
    feats, pm, dv = synthetic.make_stage_inputs(stage, 2, 3, 4 * 37, 4 * 45, 5, seed=4, channels=C)
    feats = [f[:, :, :37, :45].contiguous() for f in feats]
    dv = dv[:, :, :37, :45].contiguous()
    net, _ = _build_net(dm, sd, stage, mode)
    for hyp in (dv, dv[:, :, 0, 0].contiguous()):
        want = O.aggregate(feats, pm, hyp, mode, sd, stage)
        got = net.cost_volume(stage, [f.to(dev()) for f in feats], pm.to(dev()), hyp.to(dev()),
                              out_dtype=torch.float32).to_ncdhw().cpu()
        scale = max(want.abs().mean().item(), 1.0)
        assert (got - want).abs().max() < 2e-3 * scale
        assert (got - want).abs().mean() < 1e-5 * scale
    bf = net.cost_volume(stage, [f.to(dev()) for f in feats], pm.to(dev()), dv.to(dev()),
                         out_dtype=torch.bfloat16).to_ncdhw().cpu()
    want = O.aggregate(feats, pm, dv, mode, sd, stage)
    assert ((bf - want).abs() <= want.abs() * 2 ** -7 + 2e-3).all()


def test_half_repack_matches_cast(dm):
    x = torch.randn(2, 16, 19, 23, device=dev()) * 3
    x[0, 0, 0, 0] = 1e6                      # saturates instead of overflowing to inf
    y = dm.ops.features_to_nhwc_half(x)
    want = x.permute(0, 2, 3, 1).clamp(-65504, 65504).to(torch.float16)
    assert y.dtype == torch.float16 and y.is_contiguous() and torch.equal(y, want)
    yc = dm.ops.features_to_nhwc_half(x.contiguous(memory_format=torch.channels_last))
    assert torch.equal(yc, want)


@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("C", [8, 16, 32])
def test_half_feature_path_matches_oracle(dm, mode, C):
    """The fp16-feature kernel (the bf16 pipeline's producer) against the oracle on the SAME fp16-rounded
    features, fp32 volume out: what remains is fp16 bilinear weights (2^-12 relative) and fp32 rounding of
    the simplified coordinate arithmetic.  Ragged extents, 2 batch items, both hypothesis layouts."""
    from damvsnet_b200 import synthetic
    stage = {32: 0, 16: 1, 8: 2}[C]
    sd = synthetic.hot_path_state_dict(seed=11)
    feats, pm, dv = synthetic.make_stage_inputs(stage, 2, 4, 4 * 37, 4 * 45, 6, seed=4, channels=C)
    feats = [f[:, :, :37, :45].contiguous().half().float() for f in feats]
    dv = dv[:, :, :37, :45].contiguous()
    net, _ = _build_net(dm, sd, stage, mode)
    wnet = net.weight_net[stage].folded() if mode == "adaptive" else None
    rt = net.stage_rot_trans(pm.to(dev()))
    nhwc = [dm.ops.features_to_nhwc_half(f.to(dev())) for f in feats]
    for hyp in (dv, dv[:, :, 0, 0].contiguous()):
        want = O.aggregate(feats, pm, hyp, mode, sd, stage)
        got = dm.ops.warp_aggregate(nhwc[0], nhwc[1:], rt, hyp.to(dev()), wnet, mode, torch.float32).to_ncdhw().cpu()
        scale = max(want.abs().mean().item(), 1.0)
        assert ((got - want).abs() <= 2e-3 * want.abs() + 5e-3 * scale).all()
        assert (got - want).abs().mean() < 2e-4 * scale


def test_variance_of_identical_views_is_zero_full_size(dm):
    """Size-independent property at the BASELINE stage-1 extent: identical views under identity
    relative pose sample themselves (up to the W/(W-1) quirk) -- use a constant feature so the
    variance must vanish exactly, and the adaptive cost must be exactly zero."""
    B, C, D, H, W = 1, 32, 48, 288, 400
    f = torch.full((B, C, H, W), 0.75, device=dev())
    pm = torch.zeros(B, 3, 2, 4, 4)
    pm[:, :, 0] = torch.eye(4)
    pm[:, :, 1, :3, :3] = torch.tensor([[720.0, 0, 199.5], [0, 720.0, 143.5], [0, 0, 1]])
    dv = (425 + 10.0 * torch.arange(D, dtype=torch.float32)).view(1, D).to(dev())
    net = dm.DepthNet("variance", [32, 16, 8]).eval().to(dev())
    vol = net.cost_volume(0, [f, f, f], pm.to(dev()), dv, out_dtype=torch.float32).data
    inner = vol[:, :, :, 2:-2, 2:-2]
    assert inner.abs().max().item() < 1e-6


# ---------------------------------------------------------------- hypothesis sampling (upstream neighbour)
def test_stage_hypotheses_match_reference_fixture(dm):
    fx = golden_io.load_hypotheses()
    for name in ("stage2", "stage3"):
        H, W, scale, D = (int(v) for v in fx[name + "/meta"])
        got = dm.ops.stage_hypotheses(fx[name + "/depth"].to(dev()), fx[name + "/var"].to(dev()), D, H, W, scale).cpu()
        assert got.shape == fx[name + "/out"].shape
        torch.testing.assert_close(got, fx[name + "/out"], rtol=2e-6, atol=2e-3)
        # reference-signature entry point on full-resolution [B,1,H,W] inputs
        cur = torch.nn.functional.interpolate(fx[name + "/depth"].unsqueeze(1), [H, W], mode="bilinear", align_corners=False)
        ev = torch.nn.functional.interpolate(fx[name + "/var"].unsqueeze(1), [H, W], mode="bilinear", align_corners=False)
        full = dm.uncertainty_aware_samples(cur.to(dev()), ev.to(dev()), D, torch.float32, dev(), [2, H, W]).cpu()
        torch.testing.assert_close(full[:, :, ::3, ::3], fx[name + "/full_sub"], rtol=2e-6, atol=2e-3)
    first = dm.uncertainty_aware_samples(fx["stage1/depth_values"].to(dev()), None, 48, torch.float32, dev(), [2, 64, 96]).cpu()
    torch.testing.assert_close(first[:, :, ::16, ::16], fx["stage1/out"], rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("scale,D", [(1, 8), (2, 32), (4, 48)])
def test_stage_hypotheses_match_oracle_full_size(dm, scale, D):
    """BASELINE shapes (1152x1600): previous stage at half the output stage's pitch; monotone hypotheses."""
    g = torch.Generator().manual_seed(scale)
    H, W = 1152, 1600
    hp, wp = H // (2 * scale) if scale < 4 else H // 4, W // (2 * scale) if scale < 4 else W // 4
    depth = 450 + 400 * torch.rand(1, hp, wp, generator=g)
    var = 0.2 + 8 * torch.rand(1, hp, wp, generator=g)
    got = dm.ops.stage_hypotheses(depth.to(dev()), var.to(dev()), D, H, W, scale).cpu()
    want = O.stage_hypotheses(depth, var, D, H, W, scale)
    torch.testing.assert_close(got, want, rtol=2e-6, atol=2e-3)
    assert (got[:, 1:] >= got[:, :-1]).all()


# ---------------------------------------------------------------- conv blocks
def _rand_bn(bn, g):
    with torch.no_grad():
        bn.weight.copy_(0.8 + 0.4 * torch.rand(bn.weight.shape, generator=g))
        bn.bias.copy_(0.1 * torch.randn(bn.bias.shape, generator=g))
        bn.running_mean.copy_(0.05 * torch.randn(bn.bias.shape, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(bn.bias.shape, generator=g))


@pytest.mark.parametrize("cin,cout,stride,transposed", [
    (8, 8, 1, False), (32, 8, 1, False), (8, 16, 2, False), (16, 32, 2, False), (64, 64, 1, False),
    (64, 32, 1, True), (16, 8, 1, True)])
def test_conv_block_fp32_matches_torch(dm, cin, cout, stride, transposed):
    g = torch.Generator().manual_seed(cin * 100 + cout)
    if transposed:
        blk = dm.Deconv3d(cin, cout, stride=2, padding=1, output_padding=1)
    else:
        blk = dm.Conv3d(cin, cout, stride=stride, padding=1)
    _rand_bn(blk.bn, g)
    blk = blk.eval()
    x = torch.randn(2, cin, 6, 10, 14, generator=g)
    with torch.no_grad():
        want = torch.relu(blk.bn(blk.conv(x)))
    skip = torch.randn(want.shape, generator=g)
    blk = blk.to(dev())
    with dm.precision("fp32"):
        got = blk(x.to(dev())).cpu()
        vol = dm.G8Volume.from_ncdhw(x.to(dev()), torch.float32)
        sk = dm.G8Volume.from_ncdhw(skip.to(dev()), torch.float32)
        got_skip = blk.forward_g8(vol, skip=sk).to_ncdhw().cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max() < 2e-5 * max(1.0, want.abs().max().item())
    assert (got_skip - (want + skip)).abs().max() < 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_cost_reg_net_fp32_matches_oracle(dm, stage):
    from damvsnet_b200 import synthetic
    sd = synthetic.hot_path_state_dict(seed=2)
    cin = synthetic.STAGE_CHANNELS[stage]
    _, cr = _build_net(dm, sd, stage, "variance")
    x = torch.rand(1, cin, 8, 16, 24) * 0.5
    want = O.cost_reg_net(x, sd, stage)
    with dm.precision("fp32"):
        got = cr(x.to(dev())).cpu()
    assert got.shape == want.shape == (1, 1, 8, 16, 24)
    assert (got - want).abs().max() < 1e-4 * max(1.0, want.abs().max().item())


def test_cost_reg_net_rejects_bad_extent(dm):
    cr = dm.CostRegNet(8, 8).eval().to(dev())
    with pytest.raises(ValueError):
        cr(torch.zeros(1, 8, 8, 12, 16, device=dev()))
    with pytest.raises(ValueError):
        dm.ops.softmax_regress(torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4))  # CPU tensors: no CPU path


# ---------------------------------------------------------------- whole stage
def _check_stage(out, st, depth_max, depth_p99, conf_tol, conf_frac, depth_median=None):
    rel = _rel(out["depth"].cpu(), st["depth"])
    if depth_max is not None:
        assert rel.max().item() < depth_max, rel.max().item()
    if depth_median is not None:
        assert rel.median().item() < depth_median, rel.median().item()
    assert torch.quantile(rel.flatten(), 0.99).item() < depth_p99, torch.quantile(rel.flatten(), 0.99).item()
    bad = ((out["photometric_confidence"].cpu() - st["photometric_confidence"]).abs() > conf_tol).float().mean().item()
    assert bad <= conf_frac, bad
    assert set(out) == {"depth", "photometric_confidence", "variance", "prob_volume", "depth_values"}
    assert out["prob_volume"].shape == st["prob_volume"].shape and out["depth"].shape == st["depth"].shape


@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_depthnet_fp32_matches_reference_fixture(dm, mode, stage):
    """The headline parity claim: fp32 path, relative depth error <= 1e-4 against the reference's outputs."""
    sd, stages = golden_io.load_depthnet(mode)
    st = stages[stage]
    net, cr = _build_net(dm, sd, stage, mode)
    with dm.precision("fp32"), torch.no_grad():
        out = net(stage, [f.to(dev()) for f in st["features"]], st["proj"].to(dev()), st["depth_values"].to(dev()),
                  st["depth_values"].shape[1], cr)
    _check_stage(out, st, depth_max=1e-4, depth_p99=5e-5, conf_tol=2e-3, conf_frac=2e-3)
    assert (out["prob_volume"].cpu() - st["prob_volume"]).abs().max() < 2e-3
    vrel = _rel(out["variance"].cpu(), st["variance"], 1e-2)
    assert torch.quantile(vrel.flatten(), 0.99).item() < 1e-3


# (median, p99) of |depth - ref| / span, prob_volume max abs, confidence abs p99 -- per 2-byte pipeline
HALF_BOUNDS = {"bf16": (2.5e-3, 3e-2, 0.1, 6e-2), "fp16": (4e-4, 5e-3, 2e-2, 1e-2)}


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("mode", ["adaptive", "variance"])
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_depthnet_half_precision_within_stated_bound(dm, mode, stage, prec):
    """The reduced-precision pipelines (fp16 features; cost volume, CostRegNet weights and activations in bf16 or fp16;
    tcgen05 convolutions with fp32 accumulation), teacher-forced per stage against the reference's fp32 outputs.  The
    error is normalised by the per-pixel hypothesis span (max - min of depth_values), the natural scale of a soft-argmax
    over that span.  Bounds (HALF_BOUNDS): bf16 -- median <= 2.5e-3, p99 <= 3e-2, prob_volume max abs <= 0.1, confidence
    abs p99 <= 6e-2 (measured on B200: median 0.7-1.4e-3, p99 0.6-1.9e-2); fp16 (8x finer rounding) -- median <= 4e-4,
    p99 <= 5e-3, prob <= 2e-2, confidence p99 <= 1e-2.  There is no max-norm bound: the fixture's heads are sharpened x6
    random-init nets, where a rounding-sized logit perturbation moves probability mass between competing depth modes at
    isolated pixels (scripts/ablate_precision.py separates the contributions)."""
    sd, stages = golden_io.load_depthnet(mode)
    st = stages[stage]
    net, cr = _build_net(dm, sd, stage, mode)
    with dm.precision(prec), torch.no_grad():
        out = net(stage, [f.to(dev()) for f in st["features"]], st["proj"].to(dev()), st["depth_values"].to(dev()),
                  st["depth_values"].shape[1], cr)
    assert set(out) == {"depth", "photometric_confidence", "variance", "prob_volume", "depth_values"}
    b_med, b_p99, b_prob, b_conf = HALF_BOUNDS[prec]
    span = (st["depth_values"].max(1).values - st["depth_values"].min(1).values).clamp_min(1e-3)
    nrm = (out["depth"].cpu() - st["depth"]).abs() / span
    assert nrm.median().item() < b_med, nrm.median().item()
    assert torch.quantile(nrm.flatten(), 0.99).item() < b_p99, torch.quantile(nrm.flatten(), 0.99).item()
    assert (out["prob_volume"].cpu() - st["prob_volume"]).abs().max().item() < b_prob
    cerr = (out["photometric_confidence"].cpu() - st["photometric_confidence"]).abs()
    assert torch.quantile(cerr.flatten(), 0.99).item() < b_conf


# ---------------------------------------------------------------- other configurations of BASELINE.json
@pytest.mark.parametrize("stage,D", [(0, 64), (1, 32), (2, 8)])
def test_depthnet_batch2_seven_views_default_depths(dm, stage, D):
    """B = 2, N = 7 (Tanks-and-Temples view count), the reference's default 64/32/8 hypothesis counts
    (train.py:63), ragged-free small extent: fp32 path against the oracle."""
    from damvsnet_b200 import synthetic
    sd = synthetic.hot_path_state_dict(seed=21)
    feats, pm, dv = synthetic.make_stage_inputs(stage, 2, 7, 64, 96, D, seed=9)
    net, cr = _build_net(dm, sd, stage, "adaptive")
    want = O.depthnet_forward(stage, feats, pm, dv, sd, "adaptive")
    with dm.precision("fp32"), torch.no_grad():
        out = net(stage, [f.to(dev()) for f in feats], pm.to(dev()), dv.to(dev()), D, cr)
    rel = _rel(out["depth"].cpu(), want["depth"])
    assert rel.max().item() < 1e-4, rel.max().item()
    assert (out["prob_volume"].cpu() - want["prob_volume"]).abs().max() < 2e-3
    with dm.precision("bf16"), torch.no_grad():
        out16 = net(stage, [f.to(dev()) for f in feats], pm.to(dev()), dv.to(dev()), D, cr)
    span = (dv.max(1).values - dv.min(1).values).clamp_min(1e-3)
    nrm = (out16["depth"].cpu() - want["depth"]).abs() / span
    assert nrm.median().item() < 2.5e-3 and torch.quantile(nrm.flatten(), 0.99).item() < 3e-2


def test_runner_cascade_chains_stages_like_the_oracle(dm):
    """Three stages chained through the native hypothesis sampler (fp32 mode) against the same chain in the oracle."""
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner
    H, W, N, nds = 64, 96, 3, [16, 8, 8]
    sd = synthetic.hot_path_state_dict(seed=3)
    feats = [synthetic.make_stage_inputs(s, 1, N, H, W, nds[s], seed=6)[0] for s in range(3)]
    projs, _ = synthetic.make_cameras(1, N, H, W, seed=6)
    dvals = synthetic.make_depth_range(1, 192)
    runner = HotPathRunner(sd, device=dev())
    with dm.precision("fp32"):
        got = runner.run_cascade([[f.to(dev()) for f in fs] for fs in feats], {k: v.to(dev()) for k, v in projs.items()},
                                 dvals.to(dev()), nds, H, W)
    depth = var = None
    for s in range(3):
        h, w = H // synthetic.STAGE_SCALES[s], W // synthetic.STAGE_SCALES[s]
        if depth is None:
            dv = O.first_stage_samples(dvals, nds[s]).view(1, nds[s], 1, 1).expand(1, nds[s], h, w).contiguous()
        else:
            dv = O.stage_hypotheses(depth, var, nds[s], H, W, synthetic.STAGE_SCALES[s])
        want = O.depthnet_forward(s, feats[s], projs[f"stage{s + 1}"], dv, sd, "adaptive")
        depth, var = want["depth"], want["variance"]
        g = got[f"stage{s + 1}"]
        # not teacher forced: the stage-2/3 hypotheses hang off the previous stage's depth and variance (a square
        # root), so fp32 differences compound along the chain
        e = _rel(g["depth_values"].cpu(), dv)
        assert e.max() < (1e-6 if s == 0 else 2e-2) and e.quantile(0.99) < 2e-4
        assert _rel(g["depth"].cpu(), want["depth"]).quantile(0.99) < 1e-4 * (1 + 10 * s)
    assert set(got) == {"stage1", "stage2", "stage3", "depth", "photometric_confidence", "variance", "prob_volume", "depth_values"}
    assert got["depth"] is got["stage3"]["depth"]


def test_runner_graph_replay_matches_eager(dm):
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner, make_workload
    sd = synthetic.hot_path_state_dict(seed=3)
    runner = HotPathRunner(sd, device=dev())
    stages = make_workload(64, 96, 3, [16, 8, 8], seed=2, device=dev())
    eager = [{k: v.clone() for k, v in o.items()} for o in runner.run_device(stages)]
    for _ in range(2):
        replay = runner.run_device_graphed(stages)
    torch.cuda.synchronize()
    for a, b in zip(eager, replay):
        for k in ("depth", "photometric_confidence", "variance", "prob_volume"):
            assert torch.equal(a[k], b[k]), k


def test_runner_host_pipeline_matches_device(dm):
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner, make_workload
    sd = synthetic.hot_path_state_dict(seed=3)
    runner = HotPathRunner(sd, device=dev())
    host = make_workload(64, 96, 3, [16, 8, 8], seed=2)
    ref = runner.run_device([([f.to(dev()) for f in fs], p.to(dev()), d.to(dev())) for fs, p, d in host])
    pinned = runner.pin_stages(host)
    t1 = runner.submit_host(pinned)
    t2 = runner.submit_host(pinned)
    for t in (t1, t2):
        got = runner.collect(t)
        for a, b in zip(ref, got):
            for k in ("depth", "photometric_confidence", "variance"):
                assert torch.equal(a[k].cpu(), b[k]), k
        runner.release(t)


def test_runner_host_pipeline_fp16_channels_last_features(dm):
    """pin_stages(feature_format="nhwc_f16"): the host hands over fp16 channels_last features (what the gather kernel
    consumes); the device side is zero-copy and the result is bit-identical to the bf16 pipeline fed fp32 NCHW features
    (the repack kernel and the host conversion both round to nearest even)."""
    from damvsnet_b200 import _lib, synthetic
    from damvsnet_b200.runner import HotPathRunner, make_workload
    sd = synthetic.hot_path_state_dict(seed=3)
    runner = HotPathRunner(sd, device=dev())
    host = make_workload(64, 96, 3, [16, 8, 8], seed=2)
    with dm.precision("bf16"):
        ref = runner.run_device([([f.to(dev()) for f in fs], p.to(dev()), d.to(dev())) for fs, p, d in host])
        pinned = runner.pin_stages(host, feature_format="nhwc_f16")
        assert pinned[0][0][0].dtype == torch.float16 and pinned[0][0][0].shape == host[0][0][0].shape
        assert runner.h2d_bytes(pinned) < runner.h2d_bytes(host)
        for _ in range(2):
            n0 = _lib.launch_count()
            t = runner.submit_host(pinned)
            got = runner.collect(t)
            launches = _lib.launch_count() - n0
            for a, b in zip(ref, got):
                for k in ("depth", "photometric_confidence", "variance"):
                    assert torch.equal(a[k].cpu(), b[k]), k
            runner.release(t)
        assert launches == 3 * 14, launches          # warp + 12 conv launches (conv6 = 2) + head per stage: no repack launch
        with pytest.raises(ValueError):
            runner.pin_stages(host, feature_format="nhwc_bf16")


def test_packed_weight_cache_follows_data_writes_after_invalidate(dm):
    """writes through `.data` do not bump tensor._version; dm.invalidate_packed() (called by
    broadcast_module_state) makes the eval path re-pack.  Also: module moves invalidate on their own."""
    cr = dm.CostRegNet(8, 8).eval().to(dev())
    x = torch.randn(1, 8, 8, 16, 24, device=dev())
    with dm.precision("bf16"), torch.no_grad():
        y0 = cr(x).clone()
        cr.conv0.conv.weight.data.mul_(-1.0)            # invisible to the version counter
        cr.prob.weight.data.mul_(2.0)
        dm.invalidate_packed()
        y1 = cr(x).clone()
        assert not torch.allclose(y0, y1)
        cr.conv0.conv.weight.data.mul_(-1.0)
        cr.prob.weight.data.mul_(0.5)
        dm.invalidate_packed()
        assert torch.equal(cr(x), y0)
        with torch.no_grad():
            cr.prob.weight.mul_(2.0)                    # visible write: no invalidate needed
        assert not torch.allclose(cr(x), y0)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_depthnet_with_fused_prob_head_equals_the_two_kernel_path(dm, prec):
    """ops.set_fusion(prob_head=True): DepthNet output dict from the fused last-layer + head launch against the default
    two-kernel path, all three stages of the fixture."""
    from damvsnet_b200 import _lib, ops
    sd, stages = golden_io.load_depthnet("adaptive")
    for stage, st in enumerate(stages):
        net, cr = _build_net(dm, sd, stage, "adaptive")
        args = ([f.to(dev()) for f in st["features"]], st["proj"].to(dev()), st["depth_values"].to(dev()), st["depth_values"].shape[1], cr)
        with dm.precision(prec), torch.no_grad():
            net(stage, *args)                                              # packs the weights (launches of its own)
            n0 = _lib.launch_count()
            want = net(stage, *args)
            n1 = _lib.launch_count()
            ops.set_fusion(prob_head=True)
            try:
                got = net(stage, *args)
            finally:
                ops.set_fusion(prob_head=False)
            n2 = _lib.launch_count()
        assert (n2 - n1) == (n1 - n0) - 1                                  # one launch fewer
        assert set(got) == set(want)
        for k in ("depth", "photometric_confidence", "variance", "prob_volume"):
            tol = 1e-6 if k == "prob_volume" else 1e-5
            assert (got[k] - want[k]).abs().max().item() <= tol * max(1.0, want[k].abs().max().item()), (stage, k)


def test_tanks_and_temples_shape_seven_views_invariants(dm):
    """BASELINE.json configs[2] (1056x1920, N=7, D=48/32/8) at full size, where the oracle is too slow: size-independent
    properties of the three stages in the benchmarked bf16 mode -- probabilities sum to one, the regressed depth lies
    inside its pixel's hypothesis range, confidence in [0,1], variance >= 0, everything finite -- and agreement of the
    bf16 pipeline with the exact fp32 pipeline on the same inputs (median error normalised by the hypothesis span)."""
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner, make_workload
    sd = synthetic.hot_path_state_dict(seed=0)
    runner = HotPathRunner(sd, device=dev())
    stages = make_workload(1056, 1920, 7, [48, 32, 8], seed=2, device=dev())
    with dm.precision("bf16"):
        outs = runner.run_device(stages)
    with dm.precision("fp32"):
        ref = runner.run_stage(2, *stages[2])
    for (f, p, d), o in zip(stages, outs):
        assert len(f) == 7
        for k in ("depth", "photometric_confidence", "variance", "prob_volume"):
            assert torch.isfinite(o[k]).all(), k
        assert (o["prob_volume"].sum(1) - 1).abs().max() < 1e-4
        lo, hi = d.min(1).values, d.max(1).values
        assert (o["depth"] >= lo - 1e-3).all() and (o["depth"] <= hi + 1e-3).all()
        assert (o["photometric_confidence"] >= 0).all() and (o["photometric_confidence"] <= 1 + 1e-5).all()
        assert (o["variance"] >= 0).all()
    d3 = stages[2][2]
    span = (d3.max(1).values - d3.min(1).values).clamp_min(1e-6)
    err = ((outs[2]["depth"] - ref["depth"]).abs() / span)
    assert err.median() < 2.5e-3 and err.quantile(0.99) < 3e-2


def test_host_cascade_pipeline_matches_device_cascade(dm):
    """submit_host(cascade=...) -- features + cameras + plane-sweep range from pinned host memory, hypotheses of stages
    2/3 sampled on the device -- gives what run_cascade gives on device-resident inputs."""
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner
    H, W, N, nds = 64, 96, 3, [16, 8, 8]
    sd = synthetic.hot_path_state_dict(seed=3)
    feats = [synthetic.make_stage_inputs(s, 1, N, H, W, nds[s], seed=6)[0] for s in range(3)]
    projs, _ = synthetic.make_cameras(1, N, H, W, seed=6)
    dvals = synthetic.make_depth_range(1, 192)
    runner = HotPathRunner(sd, device=dev())
    want = runner.run_cascade([[f.to(dev()) for f in fs] for fs in feats], {k: v.to(dev()) for k, v in projs.items()},
                              dvals.to(dev()), nds, H, W)
    pinned = [([f.pin_memory() for f in feats[s]], projs[f"stage{s + 1}"].pin_memory(), None) for s in range(3)]
    for _ in range(2):      # second submit reuses the device / host buffers
        t = runner.submit_host(pinned, cascade=(dvals.pin_memory(), nds, H, W))
        got = runner.collect(t)
        for s in range(3):
            for k in ("depth", "photometric_confidence", "variance"):
                torch.testing.assert_close(got[s][k], want[f"stage{s + 1}"][k].cpu(), rtol=1e-5, atol=1e-4)
        runner.release(t)


def test_host_pipeline_teacher_forced_cascade_uploads_maps_not_hypotheses(dm):
    """submit_host with a (prev_depth, prev_variance) pair in place of a stage's hypotheses: the maps are uploaded and
    sampled on the device by the fused sampler; the result equals run_device on hypotheses sampled the same way."""
    from damvsnet_b200 import ops, synthetic
    from damvsnet_b200.runner import HotPathRunner
    H, W, N, nds = 64, 96, 3, [16, 8, 8]
    sd = synthetic.hot_path_state_dict(seed=3)
    runner = HotPathRunner(sd, device=dev())
    host = [synthetic.make_stage_inputs(s, 1, N, H, W, nds[s], seed=6) for s in range(3)]
    rng = synthetic.make_depth_range(1, 192)
    prev = [None, synthetic.make_prev_maps(1, 1, H, W, seed=1), synthetic.make_prev_maps(2, 1, H, W, seed=1)]
    dev_stages = []
    for s, (fs, p, _) in enumerate(host):
        b, _, h, w = fs[0].shape
        dv = runner.range_hypotheses(rng, nds[s], b, h, w) if s == 0 else \
            ops.stage_hypotheses(prev[s][0].to(dev()), prev[s][1].to(dev()), nds[s], H, W, H // h)
        dev_stages.append(([f.to(dev()) for f in fs], p.to(dev()), dv))
    for prec, fmt in (("fp32", "nchw_f32"), ("fp16", "nhwc_f16")):
        with dm.precision(prec):
            want = runner.run_device(dev_stages)
            pinned = runner.pin_stages([(fs, p, prev[s]) for s, (fs, p, _) in enumerate(host)], feature_format=fmt)
            assert runner.h2d_bytes(pinned) < runner.h2d_bytes(runner.pin_stages(host, feature_format=fmt))
            for _ in range(2):
                t = runner.submit_host(pinned, cascade=(rng.pin_memory(), nds, H, W))
                got = runner.collect(t)
                for s in range(3):
                    for k in ("depth", "photometric_confidence", "variance"):
                        assert torch.equal(got[s][k], want[s][k].cpu()), (prec, s, k)
                runner.release(t)


@pytest.mark.parametrize("mode", ["adaptive", "variance"])
def test_cost_volume_edge_cases_single_source_and_out_of_view(dm, mode):
    """N = 2 (one source view) and a source camera that looks away from the scene (every tap out of range: the warped
    feature is zero padding everywhere, so adaptive = (w + 1) * ref^2 and variance = E[x^2] - E[x]^2 with x_src = 0),
    on both feature paths, smallest legal image (8 x 8)."""
    from damvsnet_b200 import synthetic
    sd = synthetic.hot_path_state_dict(seed=11)
    for (H, W) in ((8 * 4, 8 * 4), (4 * 21, 4 * 33)):
        feats, pm, dv = synthetic.make_stage_inputs(0, 1, 2, H, W, 4, seed=8)
        net, _ = _build_net(dm, sd, 0, mode)
        pm_away = pm.clone()
        pm_away[:, 1, 0, :3, 3] = torch.tensor([1e6, -1e6, 0.0])             # translate the source camera far away
        for proj in (pm, pm_away):
            want = O.aggregate(feats, proj, dv, mode, sd, 0)
            for dt, tol in ((torch.float32, 2e-3), (torch.bfloat16, 2e-2)):
                got = net.cost_volume(0, [f.to(dev()) for f in feats], proj.to(dev()), dv.to(dev()), out_dtype=dt).to_ncdhw().cpu()
                scale = max(want.abs().mean().item(), 1.0)
                assert ((got - want).abs() <= tol * (want.abs() + scale)).all(), (H, W, dt)
        if mode == "variance":
            away = O.aggregate(feats, pm_away, dv, mode, sd, 0)
            ref = feats[0].unsqueeze(2).expand_as(away)
            torch.testing.assert_close(away, ref ** 2 / 2 - (ref / 2) ** 2, rtol=1e-5, atol=1e-6)
