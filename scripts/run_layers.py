"""Run a few hot-path kernels in isolation at BASELINE stage shapes (for ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import ops, synthetic
from damvsnet_b200.runner import HotPathRunner, make_workload

dev = torch.device("cuda:0")
torch.set_grad_enabled(False)   # inference path (fused kernels), as the runner uses
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = synthetic.hot_path_state_dict(seed=0)
runner = HotPathRunner(sd, device=dev)
H, W = 1152, 1600
if which in ("all", "conv"):
    # stage-3 conv0 (8->8 @ 8x1152x1600), stage-1 conv0 (32->8 @ 48x288x400), stage-2 conv11 / prob
    for stage, (C, D, h, w) in ((2, (8, 8, H, W)), (0, (32, 48, H // 4, W // 4))):
        cr = runner.cost_regularization[stage]
        vol = dm.G8Volume(torch.randn(1, C // 8, D, h, w, 8, device=dev).bfloat16())
        for _ in range(reps):
            c0 = cr.conv0.forward_g8(vol)
        x = dm.G8Volume(torch.randn(1, 2, D // 2, h // 2, w // 2, 8, device=dev).bfloat16())
        for _ in range(reps):
            y = cr.conv11.forward_g8(x, skip=c0)
        impl = ops.conv_impl_for(8, 1, 1, False)
        for _ in range(reps):
            ops.conv3d(y, cr._prob_prepared(impl), None, None, 1, 1, False, False, None, torch.float32, True, impl)
if which in ("all", "warp", "head"):
    stages = make_workload(H, W, 5, [48, 32, 8], seed=0, device=dev)
    for s, (f, p, d) in enumerate(stages):
        if which in ("all", "warp"):
            for _ in range(reps):
                vol = runner.depthnet.cost_volume(s, f, p, d)
        if which in ("all", "head"):
            logits = torch.randn(1, d.shape[1], d.shape[2], d.shape[3], device=dev)
            for _ in range(reps):
                ops.softmax_regress(logits, d)
torch.cuda.synchronize()
print("ok")
