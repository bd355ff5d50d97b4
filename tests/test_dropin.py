"""The drop-in installer against the reference's own model code (oracle/_ref staged by oracle/make_ref.py, or the live
checkout in the build container).  CPU part: construction, state_dict layout, the tensor-op side entries.  GPU part
(-m gpu): the reference's CascadeMVSNet.forward executed end to end with the hot path re-bound to this package."""
import contextlib
import io
import os
import warnings

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference code neither staged (oracle/_ref) nor mounted")


def _cascade(cas, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return cas.CascadeMVSNet(refine=False, ndepths=[48, 32, 8], depth_interals_ratio=[4, 2, 1], cr_base_chs=[8, 8, 8], **kw)


def test_install_rebinds_hot_path_and_keeps_state_dict_layout():
    import damvsnet_b200 as dm
    from damvsnet_b200 import dropin
    cas, _ = ref_loader.load()
    dropin.uninstall()
    ref_model = _cascade(cas)
    ref_keys = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    before = dm.ops.get_precision()
    try:
        patched = dropin.install()
        assert dm.ops.get_precision() == "fp32"          # the drop-in keeps the reference's arithmetic width unless asked
        assert "models.cas_mvsnet.DepthNet" in patched and "models.module.homo_warping" in patched
        model = _cascade(cas)
        assert isinstance(model.DepthNet, dm.DepthNet)
        assert all(isinstance(m, dm.CostRegNet) for m in model.cost_regularization)
        keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert keys == ref_keys                       # 580 keys, identical names and shapes
        assert len(keys) == 580
        model.load_state_dict(ref_model.state_dict(), strict=True)   # test_uni.py:224 keeps working
        # the variance variant constructs too (share_cr=True is broken in the reference itself: it passes the
        # list of stage channels as in_channels, reference models/cas_mvsnet.py:178)
        with contextlib.redirect_stdout(io.StringIO()):
            v = cas.CascadeMVSNet(agg_mode="variance")
        assert len(v.state_dict()) == 526
        dropin.install(precision="bf16")
        assert dm.ops.get_precision() == "bf16"
    finally:
        dropin.uninstall()
        dm.set_precision(before)
    assert cas.DepthNet is not dm.DepthNet


def test_weight_net_stand_alone_forward_train_and_eval_match_the_reference_class():
    """AggWeightNetVolume.forward is off the hot path (fused in DepthNet) but must be a complete stand-in: train() mode
    (batch statistics, running-buffer updates, gradients) and eval() mode against the reference's own class."""
    import damvsnet_b200 as dm
    _, rm = ref_loader.load()
    torch.manual_seed(0)
    ref = rm.AggWeightNetVolume(16).train()
    ours = dm.AggWeightNetVolume(16).train()
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.rand(2, 16, 4, 6, 8, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    yr, yo = ref(x), ours(x2)
    torch.testing.assert_close(yo, yr, rtol=1e-5, atol=1e-6)
    yr.sum().backward()
    yo.sum().backward()
    torch.testing.assert_close(x2.grad, x.grad, rtol=1e-4, atol=1e-6)
    for (k, a), (k2, b) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k == k2
        torch.testing.assert_close(b.float(), a.float(), rtol=1e-5, atol=1e-6)
    ref.eval()
    ours.eval()
    with torch.no_grad():
        torch.testing.assert_close(ours(x.detach()), ref(x.detach()), rtol=1e-5, atol=1e-6)


def test_uncertainty_samples_keep_the_undetach_gradient_path():
    """grad_method != "detach" (reference models/cas_mvsnet.py:236-243): the hypothesis samples are differentiable in
    the previous stage's depth and variance."""
    import damvsnet_b200 as dm
    _, rm = ref_loader.load()
    g = torch.Generator().manual_seed(1)
    cd = (500 + 50 * torch.rand(1, 1, 6, 8, generator=g)).requires_grad_(True)
    ev = (5 + 3 * torch.rand(1, 1, 6, 8, generator=g)).requires_grad_(True)
    cd2, ev2 = cd.detach().clone().requires_grad_(True), ev.detach().clone().requires_grad_(True)
    a = rm.uncertainty_aware_samples(cd, ev, 8, torch.float32, "cpu", [1, 6, 8])
    b = dm.uncertainty_aware_samples(cd2, ev2, 8, torch.float32, "cpu", [1, 6, 8])
    torch.testing.assert_close(b, a, rtol=1e-6, atol=2e-4)
    w = torch.rand(a.shape, generator=g)
    (a * w).sum().backward()
    (b * w).sum().backward()
    torch.testing.assert_close(cd2.grad, cd.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ev2.grad, ev.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "fp16", "bf16"])
def test_reference_cascade_forward_runs_on_the_native_hot_path(prec):
    """The reference's CascadeMVSNet.forward (models/cas_mvsnet.py:190-319), images in -> output dict, executed on the
    GPU twice: as is (PyTorch ops, strict fp32), and with dropin.install() (its FPN / GeoFeatureFusion, this package's
    DepthNet / CostRegNet / samplers).  Same weights (BN-calibrated by two train-mode passes of the reference)."""
    import damvsnet_b200 as dm
    from damvsnet_b200 import _lib, dropin, synthetic
    cas, _ = ref_loader.load()
    dev = torch.device("cuda:0")
    H, W, N = 128, 160, 3
    dropin.uninstall()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        ref_model = _cascade(cas).to(dev)
        projs, intr = synthetic.make_cameras(1, N, H, W, seed=0)
        projs = {k: v.to(dev) for k, v in projs.items()}
        intr = {k: v.to(dev) for k, v in intr.items()}
        dvals = synthetic.make_depth_range(1, 192).to(dev)
        for m in ref_model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = None
        ref_model.train()
        with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            for s in range(2):
                ref_model(synthetic.make_images(1, N, H, W, seed=10 + s).to(dev), projs, dvals, intr)
        ref_model.eval()
        imgs = synthetic.make_images(1, N, H, W, seed=0).to(dev)
        with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            want = ref_model(imgs, projs, dvals, intr)
        n0 = _lib.launch_count()
        dropin.install(precision=prec)
        try:
            ours = _cascade(cas)
            ours.load_state_dict(ref_model.state_dict(), strict=True)
            ours = ours.to(dev).eval()
            with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                warnings.simplefilter("ignore")
                got = ours(imgs, projs, dvals, intr)
        finally:
            dropin.uninstall()
            dm.set_precision("fp32")
        assert _lib.launch_count() - n0 >= 3 * 14, "the native kernels did not run"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert set(got) == set(want) == {"stage1", "stage2", "stage3", "depth", "photometric_confidence", "variance", "prob_volume",
                                     "depth_values"}
    # stage 1 is teacher-forced by construction (same FPN features, same plane-sweep hypotheses)
    rel1 = ((got["stage1"]["depth"] - want["stage1"]["depth"]).abs() / want["stage1"]["depth"].abs())
    assert rel1.max().item() <= (1e-4 if prec == "fp32" else 5e-3), rel1.max().item()
    # later stages hang off the previous stage's depth AND variance (a square root) through the samplers and through
    # GeoFeatureFusion, so differences compound: quantile bounds
    for k in ("stage2", "stage3"):
        rel = ((got[k]["depth"] - want[k]["depth"]).abs() / want[k]["depth"].abs()).flatten()
        assert rel.quantile(0.99).item() <= (2e-3 if prec == "fp32" else 2e-2), (k, rel.quantile(0.99).item())
        assert got[k]["prob_volume"].shape == want[k]["prob_volume"].shape
        assert torch.isfinite(got[k]["depth"]).all()


@pytest.mark.gpu
def test_config0_cascade_512x640_five_views_against_the_reference_on_cpu():
    """BASELINE.json configs[0]: CascadeMVSNet forward, random init, DTU-train shape 512x640, N = 5, D = 48/32/8, batch 1,
    fp32 -- the reference's own CPU-runnable case.  The reference runs on the CPU as it is; the same class with the hot
    path re-bound by dropin.install("fp32") runs on the GPU (its FPN / GeoFeatureFusion in PyTorch on the GPU).  Stage 1
    is teacher-forced up to the fp32 noise of the PyTorch feature extractor on two devices; stages 2/3 compound it
    through depth, variance and GeoFeatureFusion, hence quantile bounds."""
    import damvsnet_b200 as dm
    from damvsnet_b200 import dropin, synthetic
    cas, _ = ref_loader.load()
    dev = torch.device("cuda:0")
    H, W, N = 512, 640, 5
    dropin.uninstall()
    torch.manual_seed(0)
    ref_model = _cascade(cas)
    projs, intr = synthetic.make_cameras(1, N, H, W, seed=0)
    dvals = synthetic.make_depth_range(1, 192)
    for m in ref_model.modules():                       # BN calibration: one cumulative-average train pass
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.momentum = None
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref_model = ref_model.to(dev).train()
        with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            ref_model(synthetic.make_images(1, N, H, W, seed=11).to(dev), {k: v.to(dev) for k, v in projs.items()}, dvals.to(dev),
                      {k: v.to(dev) for k, v in intr.items()})
        ref_model = ref_model.cpu().eval()
        imgs = synthetic.make_images(1, N, H, W, seed=0)
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            want = ref_model(imgs, projs, dvals, intr)                      # the reference, on the CPU, untouched
        dropin.install(precision="fp32")
        try:
            ours = _cascade(cas)
            ours.load_state_dict(ref_model.state_dict(), strict=True)
            ours = ours.to(dev).eval()
            with torch.no_grad(), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                warnings.simplefilter("ignore")
                got = ours(imgs.to(dev), {k: v.to(dev) for k, v in projs.items()}, dvals.to(dev), {k: v.to(dev) for k, v in intr.items()})
        finally:
            dropin.uninstall()
            dm.set_precision("fp32")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert got["stage1"]["depth"].shape == (1, 128, 160) and got["depth"].shape == (1, 512, 640)
    rel1 = ((got["stage1"]["depth"].cpu() - want["stage1"]["depth"]).abs() / want["stage1"]["depth"].abs()).flatten()
    assert rel1.quantile(0.999).item() <= 1e-4 and rel1.max().item() <= 1e-3, (rel1.quantile(0.999).item(), rel1.max().item())
    for k in ("stage2", "stage3"):
        rel = ((got[k]["depth"].cpu() - want[k]["depth"]).abs() / want[k]["depth"].abs()).flatten()
        assert rel[:: max(1, rel.numel() // 2 ** 22)].quantile(0.99).item() <= 2e-3, (k, rel.quantile(0.99).item())
        assert (got[k]["prob_volume"].sum(1) - 1).abs().max().item() < 1e-4
