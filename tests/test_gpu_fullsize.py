"""Full-size parity: the CUDA path against the CPU oracle at BASELINE.json's own shapes.

configs[1]  DTU-test 1152x1600, N=5, D=48/32/8: all three stages, teacher-forced, BOTH precision modes.
configs[2]  Tanks-and-Temples 1056x1920, N=7: stage 3 (the largest grid: 2.03 M pixels x 8 hypotheses x 7 views).

Grid-size bugs (tile counts beyond the resident CTAs, 32-bit offsets at 14.7 M voxels x C channels) only show at these
sizes; the fixtures of test_gpu_parity.py are 64x96.  The oracle (oracle/damvs_oracle.py, pinned to the reference's
own outputs by tests/test_oracle_golden.py) runs once per module on the host cores (~1 min on the GPU box) and is
shared by the fp32 and bf16 checks.  The net is the BN-calibrated random-init net of the reference-generated fixture
with its x6 head sharpening undone (SURVEY.md H7's setting).  When the staged reference (oracle/_ref) travelled with
the tree, the reference's OWN DepthNet is additionally run on the GPU in strict fp32 and must agree with both.

Tolerances: fp32 mode -- relative depth error max <= 1e-4 (north star).  Reduced-precision pipelines (fp16 features; cost
volume, weights and activations in fp16 or bf16; tcgen05 convolutions) -- the stated bounds of DESIGN.md section 5
(FULL_BOUNDS below), measured per component by scripts/ablate_precision.py.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import damvs_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402
from tests import golden_io  # noqa: E402
from tests.parity_metrics import stage_errors  # noqa: E402

DEV = "cuda:0"


def calibrated_state_dict():
    sharp, _ = golden_io.load_depthnet("adaptive")
    return {k: (v / 6.0 if k.endswith("prob.weight") else v.clone()) for k, v in sharp.items()}


@pytest.fixture(scope="module")
def dm():
    import damvsnet_b200 as dm
    from damvsnet_b200 import _lib
    _lib.check(_lib.load().damvs_check_device(0))
    return dm


@pytest.fixture(scope="module")
def dtu(dm):
    """(state_dict, host stages, oracle outputs) at 1152x1600, N=5, D=48/32/8."""
    import os
    from damvsnet_b200.runner import make_workload
    torch.set_num_threads(os.cpu_count() or 1)
    sd = calibrated_state_dict()
    stages = make_workload(1152, 1600, 5, [48, 32, 8], seed=0)
    want = [O.depthnet_forward(s, f, p, d, sd, "adaptive") for s, (f, p, d) in enumerate(stages)]
    for w in want:
        for k in ("depth", "prob_volume", "photometric_confidence", "variance"):
            assert torch.isfinite(w[k]).all(), k
    return sd, stages, want


def _run(dm, sd, stages, prec, only=None, features="auto"):
    from damvsnet_b200.runner import HotPathRunner
    runner = HotPathRunner(sd, device=DEV)
    outs = []
    with dm.precision(prec, features=features):
        for s, (f, p, d) in enumerate(stages):
            if only is not None and s != only:
                outs.append(None)
                continue
            outs.append(runner.run_stage(s, [x.to(DEV) for x in f], p.to(DEV), d.to(DEV)))
    torch.cuda.synchronize()
    return outs


def _frac_above(out, ref, thr=1e-4):
    rel = (out["depth"].cpu().double() - ref["depth"].cpu().double()).abs() / ref["depth"].cpu().double().abs()
    return (rel > thr).double().mean().item()


def test_dtu_full_size_fp32_matches_oracle(dm, dtu):
    """fp32 pipeline against the fp32 CPU oracle: relative depth error <= 1e-4 on >= 99.998 % of the pixels and
    <= 2e-4 everywhere.  Two independent fp32 implementations are being compared: measured on B200, ONE stage-1 pixel of
    115 200 sits at 1.1e-4, and the float64 yardstick (next test) shows 8.0e-5 of that is the reference side's own
    rounding noise and 3.1e-5 this package's."""
    sd, stages, want = dtu
    outs = _run(dm, sd, stages, "fp32")
    for s, (o, w) in enumerate(zip(outs, want)):
        e = stage_errors(o, w, stages[s][2])
        assert e["depth_rel_max"] <= 2e-4, (s, e)
        assert e["depth_rel_p99"] <= 5e-5, (s, e)
        assert _frac_above(o, w) <= 2e-5, (s, _frac_above(o, w), e)
        assert e["prob_max"] <= 2e-3, (s, e)
        assert e["conf_frac_gt_2e-3"] <= 2e-3, (s, e)
        assert e["var_rel_p99"] <= 1e-3, (s, e)
        assert tuple(o["prob_volume"].shape) == tuple(w["prob_volume"].shape)


# per reduced-precision pipeline: (depth rel median, p99, max), prob_volume max abs, confidence abs p99.  Stage 1 sweeps
# the whole 506 mm range with a peaked (median peak probability 0.5) multi-modal volume, so it carries the bound; stages
# 2/3 sweep 24 / 6 mm around ~600 mm and sit two orders of magnitude below it (scripts/ablate_precision.py).
FULL_BOUNDS = {"fp16": (4e-4, 2.5e-3, 1e-2, 2e-2, 5e-2), "bf16": (2.5e-3, 2e-2, 6e-2, 0.15, 0.4)}


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_dtu_full_size_half_precision_within_stated_bound(dm, dtu, prec):
    sd, stages, want = dtu
    outs = _run(dm, sd, stages, prec)
    b_med, b_p99, b_max, b_prob, b_conf = FULL_BOUNDS[prec]
    for s, (o, w) in enumerate(zip(outs, want)):
        e = stage_errors(o, w, stages[s][2])
        assert e["depth_rel_median"] <= b_med, (s, e)
        assert e["depth_rel_p99"] <= b_p99, (s, e)
        assert e["depth_rel_max"] <= b_max, (s, e)
        assert e["prob_max"] <= b_prob, (s, e)
        assert e["conf_p99"] <= b_conf, (s, e)
        if s > 0:        # narrow sweeps: far inside SURVEY.md H7's bound (p99 <= 1e-3, max <= 5e-3)
            assert e["depth_rel_p99"] <= 1e-4 and e["depth_rel_max"] <= 5e-4, (s, e)
        for k in ("depth", "photometric_confidence", "variance", "prob_volume"):
            assert torch.isfinite(o[k]).all(), (s, k)


def test_dtu_full_size_fp16_convs_with_fp32_features_meet_the_survey_bound(dm, dtu):
    """set_precision("fp16", features="fp32"): the gather stays in fp32, the cost volume and CostRegNet run in fp16 on the
    tensor cores.  This is the configuration SURVEY.md H7's bound was written for (reduced precision on the convolutions
    only): per stage depth rel p99 <= 1e-3, max <= 5e-3, confidence abs p99 <= 1e-2 -- met at full size on the peaked net."""
    sd, stages, want = dtu
    outs = _run(dm, sd, stages, "fp16", features="fp32")
    for s, (o, w) in enumerate(zip(outs, want)):
        e = stage_errors(o, w, stages[s][2])
        assert e["depth_rel_p99"] <= 1e-3 and e["depth_rel_max"] <= 5e-3 and e["conf_p99"] <= 1e-2 and e["prob_max"] <= 2e-2, (s, e)


@pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref was not staged with this tree")
def test_dtu_full_size_reference_itself_on_gpu_agrees(dm, dtu):
    """The reference's own DepthNet (unmodified, nn.Conv3d / F.grid_sample) on the same GPU in strict fp32 against the
    oracle and against this package's fp32 mode."""
    import warnings
    sd, stages, want = dtu
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        depthnet, crs = ref_loader.build_hot_path(sd, "adaptive", device=DEV)
        dev_stages = [([x.to(DEV) for x in f], p.to(DEV), d.to(DEV)) for f, p, d in stages]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = ref_loader.hot_path_forward(depthnet, crs, dev_stages)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    ours = _run(dm, sd, stages, "fp32")
    for s in range(3):
        e_oracle = stage_errors(ref[s], want[s], stages[s][2])
        assert e_oracle["depth_rel_max"] <= 2e-4 and _frac_above(ref[s], want[s]) <= 2e-5, ("reference-on-GPU vs oracle", s, e_oracle)
        e_ours = stage_errors(ours[s], {k: v for k, v in ref[s].items()}, stages[s][2])
        assert e_ours["depth_rel_max"] <= 2e-4 and _frac_above(ours[s], ref[s]) <= 2e-5, ("ours vs reference-on-GPU", s, e_ours)
        assert e_ours["prob_max"] <= 2e-3, (s, e_ours)


@pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref was not staged with this tree")
def test_dtu_full_size_fp32_against_float64_yardstick(dm, dtu):
    """The north star's fp32 claim, made exact: the reference's DepthNet evaluated in FLOAT64 on the GPU is the truth;
    this package's fp32 pipeline must be within 1e-4 relative depth of it EVERYWHERE, and no further from it than the
    reference's own fp32 path is (measured: stage 1 max 6.4e-5 ours vs 8.0e-5 reference; stages 2/3 < 1e-6 both)."""
    import warnings
    sd, stages, _ = dtu
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        d32, c32 = ref_loader.build_hot_path(sd, "adaptive", device=DEV)
        d64, c64 = ref_loader.build_hot_path(sd, "adaptive", device=DEV)
        d64, c64 = d64.double(), c64.double()
        ours = _run(dm, sd, stages, "fp32")
        for s, (f, p, d) in enumerate(stages):
            f, p, d = [x.to(DEV) for x in f], p.to(DEV), d.to(DEV)
            with warnings.catch_warnings(), torch.no_grad():
                warnings.simplefilter("ignore")
                truth = ref_loader.float64_stage_forward(d64, c64, s, f, p, d)
                ref = d32(s, list(f), p, d, d.shape[1], c32[s])
            t = truth["depth"]
            e_ours = ((ours[s]["depth"].double() - t).abs() / t.abs())
            e_ref = ((ref["depth"].double() - t).abs() / t.abs())
            assert e_ours.max().item() <= 1e-4, (s, e_ours.max().item())
            assert e_ours.max().item() <= 1.5 * e_ref.max().item() + 1e-6, (s, e_ours.max().item(), e_ref.max().item())
            assert e_ours.median().item() <= 1.5 * e_ref.median().item() + 1e-7, (s, e_ours.median().item(), e_ref.median().item())
            assert (ours[s]["prob_volume"].double() - truth["prob_volume"]).abs().max().item() <= 1e-3
            del truth, ref
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("stage,D", [(0, 48), (2, 8)])
def test_tnt_full_size_seven_views_matches_oracle(dm, stage, D):
    """BASELINE.json configs[2] shape (1056x1920, N=7): stage 1 (C=32, D=48, six source views) and stage 3 (the largest
    grid: 2.03 M pixels x 8 hypotheses), every precision mode against the oracle."""
    from damvsnet_b200 import synthetic
    sd = calibrated_state_dict()
    f, p, d = synthetic.make_stage_inputs(stage, 1, 7, 1056, 1920, D, seed=2)
    want = O.depthnet_forward(stage, f, p, d, sd, "adaptive")
    stages = [(None, None, None)] * 3
    stages[stage] = (f, p, d)
    o32 = _run(dm, sd, stages, "fp32", only=stage)[stage]
    e = stage_errors(o32, want, d)
    assert e["depth_rel_max"] <= 2e-4 and e["depth_rel_p99"] <= 5e-5 and _frac_above(o32, want) <= 2e-5 and e["prob_max"] <= 2e-3, e
    for prec in ("fp16", "bf16"):
        o16 = _run(dm, sd, stages, prec, only=stage)[stage]
        e = stage_errors(o16, want, d)
        if stage == 0:
            b_med, b_p99, b_max, b_prob, _ = FULL_BOUNDS[prec]
            assert e["depth_rel_median"] <= b_med and e["depth_rel_p99"] <= b_p99 and e["depth_rel_max"] <= b_max and e["prob_max"] <= b_prob, (prec, e)
        else:
            assert e["depth_rel_p99"] <= 1e-4 and e["depth_rel_max"] <= 5e-4 and e["prob_max"] <= 2e-2, (prec, e)


def test_dtu_full_size_variance_aggregation_stage3(dm):
    """agg_mode="variance" (reference models/cas_mvsnet.py:34-37, 61-63, 84-85) at the full DTU-test stage-3 grid
    (1152x1600, N=5, D=8), every precision mode against the oracle."""
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import HotPathRunner
    sd = {k: v for k, v in calibrated_state_dict().items() if not k.startswith("DepthNet.")}
    f, p, d = synthetic.make_stage_inputs(2, 1, 5, 1152, 1600, 8, seed=4)
    want = O.depthnet_forward(2, f, p, d, sd, "variance")
    runner = HotPathRunner(sd, mode="variance", device=DEV)
    for prec in ("fp32", "fp16", "bf16"):
        with dm.precision(prec):
            o = runner.run_stage(2, [x.to(DEV) for x in f], p.to(DEV), d.to(DEV))
        e = stage_errors(o, want, d)
        if prec == "fp32":
            assert e["depth_rel_max"] <= 2e-4 and e["depth_rel_p99"] <= 5e-5 and e["prob_max"] <= 2e-3, e
        else:
            assert e["depth_rel_p99"] <= 1e-4 and e["depth_rel_max"] <= 5e-4 and e["prob_max"] <= 3e-2, (prec, e)
