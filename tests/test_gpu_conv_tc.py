"""tcgen05 implicit-GEMM conv blocks against torch fp32 convolutions on bf16- / fp16-representable data.

Inputs and weights are rounded to the volumes' 2-byte type first, so the only differences left are the fp32
accumulation order and the final rounding of the output (relative 2^-8 for bf16, 2^-11 for fp16)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


HALF = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _bf(t, dt=torch.bfloat16):
    return t.to(dt).float()


CASES = [
    # cin, cout, stride, transposed, (D, H, W)
    (8, 8, 1, False, (8, 16, 24)),
    (8, 8, 1, False, (5, 19, 67)),      # ragged: several tiles, partial last tile
    (16, 8, 1, False, (4, 9, 33)),
    (32, 8, 1, False, (6, 18, 40)),
    (16, 16, 1, False, (4, 12, 36)),
    (32, 32, 1, False, (3, 10, 31)),
    (64, 64, 1, False, (2, 9, 20)),     # split into two N=32 launches
    (8, 16, 2, False, (8, 16, 24)),
    (8, 16, 2, False, (6, 22, 70)),
    (16, 32, 2, False, (4, 12, 36)),
    (32, 64, 2, False, (4, 8, 34)),
    (64, 32, 1, True, (2, 5, 9)),
    (32, 16, 1, True, (3, 9, 33)),
    (16, 8, 1, True, (4, 11, 40)),
]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("cin,cout,stride,transposed,ext", CASES)
def test_tc_conv_block_matches_torch(cin, cout, stride, transposed, ext, prec):
    import damvsnet_b200 as dm
    hd = HALF[prec]
    g = torch.Generator().manual_seed(cin * 1000 + cout * 10 + stride + (5 if transposed else 0))
    if transposed:
        blk = dm.Deconv3d(cin, cout, stride=2, padding=1, output_padding=1)
    else:
        blk = dm.Conv3d(cin, cout, stride=stride, padding=1)
    with torch.no_grad():
        blk.conv.weight.copy_(_bf(blk.conv.weight, hd))
        blk.bn.weight.copy_(0.8 + 0.4 * torch.rand(cout, generator=g))
        blk.bn.bias.copy_(0.1 * torch.randn(cout, generator=g))
        blk.bn.running_mean.copy_(0.05 * torch.randn(cout, generator=g))
        blk.bn.running_var.copy_(0.5 + torch.rand(cout, generator=g))
    blk = blk.eval()
    B = 2
    x = _bf(torch.randn(B, cin, *ext, generator=g), hd)
    with torch.no_grad():
        want = torch.relu(blk.bn(blk.conv(x)))
    skip = _bf(torch.randn(want.shape, generator=g), hd)
    blk = blk.to(dev())
    with dm.precision(prec, "tcgen05"), torch.no_grad():      # inference: BatchNorm / ReLU / skip fused into the conv epilogue
        vol = dm.G8Volume.from_ncdhw(x.to(dev()), hd)
        got = blk.forward_g8(vol).to_ncdhw().cpu()
        sk = dm.G8Volume.from_ncdhw(skip.to(dev()), hd)
        got_skip = blk.forward_g8(vol, skip=sk).to_ncdhw().cpu()
    if prec == "bf16":
        # with gradients enabled the same block runs as raw conv -> BatchNorm/ReLU/skip kernels (autograd.ConvBlockFn)
        with dm.precision(prec, "tcgen05"):
            got_tape = blk.forward_g8(vol, skip=sk).to_ncdhw().detach().cpu()
        assert (got_tape - got_skip).abs().max() <= 2 ** -6 * got_skip.abs().max() + 1e-2
    torch.cuda.synchronize()
    assert got.shape == want.shape
    tol = 2 ** -7 if prec == "bf16" else 2 ** -10
    err = (got - want).abs()
    floor = 2e-3 if prec == "bf16" else 3e-4     # fp32 accumulation-order noise on O(1) sums of up to 1728 products
    assert (err <= tol * want.abs() + floor).all(), (err.max().item(), (err / want.abs().clamp_min(1e-2)).max().item())
    ws = want + skip
    err = (got_skip - ws).abs()
    assert (err <= tol * ws.abs() + tol * want.abs() + 2 * floor).all(), err.max().item()


@pytest.mark.parametrize("shape", [
    (2, 8, 13, 45),       # several tiles, ragged in both directions
    (1, 5, 7, 33),        # odd depth: the last plane pair is half padding
    (1, 1, 6, 30),        # a single plane, exactly one tile
    (1, 2, 3, 5),         # smaller than a tile
    (1, 48, 20, 70),      # stage-1 depth: the accumulator ring wraps many times
    (1, 5, 120, 600),     # more tiles than resident CTAs, odd depth: ring position carried across tiles
    (2, 8, 126, 330),
])
@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_tc_prob_conv_plain_output(shape, prec):
    import damvsnet_b200 as dm
    from damvsnet_b200 import ops
    hd = HALF[prec]
    g = torch.Generator().manual_seed(9)
    w = _bf(torch.randn(1, 8, 3, 3, 3, generator=g) * 0.2, hd)
    x = _bf(torch.randn(shape[0], 8, *shape[1:], generator=g), hd)
    want = F.conv3d(x, w, None, padding=1).squeeze(1)
    packed = ops.conv3d_pack_weight(w.to(dev()), 8, 1, False, ops.CONV_TCGEN05, 1, hd)
    vol = dm.G8Volume.from_ncdhw(x.to(dev()), hd)
    got = ops.conv3d(vol, packed, None, None, 1, 1, False, False, None, torch.float32, True, ops.CONV_TCGEN05).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max() < 1e-4 * max(1.0, want.abs().max().item())


def test_packed_weights_of_the_other_half_type_are_rejected():
    """A blob packed for bf16 volumes must not be silently consumed as fp16 (same width, different bits)."""
    import damvsnet_b200 as dm
    from damvsnet_b200 import ops
    cr = dm.CostRegNet(8, 8).eval().to(dev())
    x = torch.rand(1, 8, 8, 16, 24, device=dev())
    with torch.no_grad():
        with dm.precision("bf16"):
            a = cr(x)
        with dm.precision("fp16"):
            b = cr(x)                       # re-packs: the cache is keyed on the volume type
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    assert (a - b).abs().max() < 0.05 * max(1.0, a.abs().max().item())
    assert len({k[-1] for k in cr.conv0._packed}) == 2


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_cost_reg_net_tc_vs_direct_bf16(stage, prec):
    """Whole U-Net: tcgen05 path against the direct fp32-accumulate kernels on the same 2-byte storage."""
    import damvsnet_b200 as dm
    from damvsnet_b200 import synthetic
    sd = synthetic.hot_path_state_dict(seed=2)
    cin = synthetic.STAGE_CHANNELS[stage]
    cr = dm.CostRegNet(cin, 8).eval()
    pre = f"cost_regularization.{stage}."
    cr.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    cr = cr.to(dev())
    x = (torch.rand(1, cin, 8, 24, 40) * 0.5).to(dev())
    with dm.precision(prec, "direct"), torch.no_grad():
        a = cr(x).cpu()
    with dm.precision(prec, "tcgen05"), torch.no_grad():
        b = cr(x).cpu()
    scale = a.abs().max().item()
    k = 1.0 if prec == "bf16" else 0.125
    assert (a - b).abs().max().item() < 3e-2 * k * max(scale, 1.0), ((a - b).abs().max().item(), scale)
    assert (a - b).abs().mean().item() < 3e-3 * k * max(scale, 1.0)


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [
    (1, 8, 13, 45),        # ragged tiles, one plane pair group
    (2, 32, 20, 70),       # batch 2, stage-2 depth
    (1, 48, 37, 61),       # stage-1 depth: the accumulator ring wraps many times per tile
    (1, 8, 126, 330),      # more tiles than resident CTAs: the logits buffer is reused tile after tile
    (1, 64, 9, 33),        # the largest fused depth
])
def test_fused_prob_head_matches_the_two_kernel_path(shape, prec):
    """damvs_prob_head_fwd (prob convolution + softmax / regression / confidence / variance in one launch, logits in
    shared memory) against damvs_conv3d_fwd + damvs_softmax_regress_fwd on the same volume: same MMAs, same head
    arithmetic in the same order."""
    import damvsnet_b200 as dm
    from damvsnet_b200 import ops
    hd = HALF[prec]
    b, d, h, w = shape
    g = torch.Generator().manual_seed(d * 7 + h)
    wgt = _bf(torch.randn(1, 8, 3, 3, 3, generator=g) * 0.6, hd)
    x = _bf(torch.randn(b, 8, d, h, w, generator=g), hd)
    dv = (425 + 2.65 * torch.arange(d, dtype=torch.float32).view(1, d, 1, 1) + torch.rand(b, 1, h, w, generator=g)).expand(b, d, h, w).contiguous()
    packed = ops.conv3d_pack_weight(wgt.to(dev()), 8, 1, False, ops.CONV_TCGEN05, 1, hd)
    vol = dm.G8Volume.from_ncdhw(x.to(dev()), hd)
    assert ops.prob_head_supported(vol, ops.CONV_TCGEN05)
    logits = ops.conv3d(vol, packed, None, None, 1, 1, False, False, None, torch.float32, True, ops.CONV_TCGEN05)
    want = ops.softmax_regress(logits, dv.to(dev()))
    got = ops.prob_head(vol, packed, dv.to(dev()), ops.CONV_TCGEN05)
    torch.cuda.synchronize()
    for name, a, b_ in zip(("prob", "depth", "conf", "var"), got, want):
        assert a.shape == b_.shape, name
        assert torch.isfinite(a).all(), name
        tol = 1e-6 if name == "prob" else 1e-5
        assert (a - b_).abs().max().item() <= tol * max(1.0, b_.abs().max().item()), (name, (a - b_).abs().max().item())
    # and against torch on the CPU (fp32 conv of the same rounded operands + the oracle's head)
    import torch.nn.functional as F
    from oracle import damvs_oracle as O
    ref = O.regress_head(F.conv3d(x, wgt, None, padding=1).squeeze(1), dv)
    assert (got[0].cpu() - ref["prob_volume"]).abs().max() < 1e-4
    assert ((got[1].cpu() - ref["depth"]).abs() / ref["depth"].abs()).max() < 1e-4


def test_fused_prob_head_falls_back_where_it_does_not_apply():
    import damvsnet_b200 as dm
    from damvsnet_b200 import ops
    for d, ok in ((8, True), (6, False), (72, False)):        # D % 8 != 0 and D > 64 keep the two-kernel path
        vol = dm.G8Volume(torch.zeros(1, 1, d, 8, 32, 8, device=dev(), dtype=torch.float16))
        assert ops.prob_head_supported(vol, ops.CONV_TCGEN05) == ok, d
    vol32 = dm.G8Volume(torch.zeros(1, 1, 8, 8, 32, 8, device=dev(), dtype=torch.float32))
    assert not ops.prob_head_supported(vol32, ops.CONV_DIRECT)
