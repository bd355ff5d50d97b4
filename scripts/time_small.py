"""Device time of the coarse-level CostRegNet layers (graph-captured) -- development aid."""
import sys, torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import ops, synthetic
from damvsnet_b200.runner import HotPathRunner
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
H, W = 1152, 1600
def timeit(fn, n=20):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n // 10): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (n // 10 * 10) * 1e3
for stage, (D, h, w) in ((2, (8, H, W)), (1, (32, H // 2, W // 2)), (0, (48, H // 4, W // 4))):
    cr = runner.cost_regularization[stage]
    mk = lambda c, s: dm.G8Volume(torch.randn(1, c // 8, D // s, h // s, w // s, 8, device=dev).bfloat16())
    x2, x4, x8 = mk(16, 2), mk(32, 4), mk(64, 8)
    c4 = mk(32, 4); c2 = mk(16, 2)
    t = [timeit(lambda: cr.conv3.forward_g8(x2)), timeit(lambda: cr.conv4.forward_g8(x4)), timeit(lambda: cr.conv5.forward_g8(x4)),
         timeit(lambda: cr.conv6.forward_g8(x8)), timeit(lambda: cr.conv7.forward_g8(x8, skip=c4)), timeit(lambda: cr.conv9.forward_g8(x4, skip=c2))]
    print(f"stage{stage+1}: conv3 {t[0]:6.1f} conv4 {t[1]:6.1f} conv5 {t[2]:6.1f} conv6 {t[3]:6.1f} conv7 {t[4]:6.1f} conv9 {t[5]:6.1f}  sum {sum(t):6.1f} us")
