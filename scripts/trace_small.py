"""Per-CTA timelines of the coarse CostRegNet layers (needs a DAMVS_TC_TRACE_BUILD=1 build and DAMVS_TC_TRACE=1) -- development aid."""
import sys, torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import synthetic
from damvsnet_b200.runner import HotPathRunner
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
H, W = 1152, 1600
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
D, h, w = {2: (8, H, W), 1: (32, H // 2, W // 2), 0: (48, H // 4, W // 4)}[stage]
cr = runner.cost_regularization[stage]
mk = lambda c, s: dm.G8Volume(torch.randn(1, c // 8, D // s, h // s, w // s, 8, device=dev).bfloat16())
x2, x4, x8, c4, c2 = mk(16, 2), mk(32, 4), mk(64, 8), mk(32, 4), mk(16, 2)
for name, fn in (("conv3", lambda: cr.conv3.forward_g8(x2)), ("conv4", lambda: cr.conv4.forward_g8(x4)), ("conv5", lambda: cr.conv5.forward_g8(x4)),
                 ("conv6", lambda: cr.conv6.forward_g8(x8)), ("conv7", lambda: cr.conv7.forward_g8(x8, skip=c4)), ("conv9", lambda: cr.conv9.forward_g8(x4, skip=c2))):
    fn(); torch.cuda.synchronize()          # warm (weights packed, module loaded)
    print(f"== {name}", file=sys.stderr, flush=True)
    fn(); torch.cuda.synchronize()
