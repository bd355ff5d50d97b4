"""Device time of the big CostRegNet layers in isolation (CUDA events, 20 reps) -- development aid."""
import sys, torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import ops, synthetic
from damvsnet_b200.runner import HotPathRunner
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
H, W = 1152, 1600
def timeit(fn, n=20):
    """Device time per call: the call is captured in a CUDA graph (10 copies) so host launch overhead
    (descriptor encode, Python) does not bound the measurement."""
    with torch.no_grad():
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
for stage, (C, D, h, w) in ((2, (8, 8, H, W)), (1, (16, 32, H // 2, W // 2)), (0, (32, 48, H // 4, W // 4))):
    cr = runner.cost_regularization[stage]
    vol = dm.G8Volume(torch.randn(1, C // 8, D, h, w, 8, device=dev).bfloat16())
    c0 = cr.conv0.forward_g8(vol)
    c1 = cr.conv1.forward_g8(c0)
    x = dm.G8Volume(torch.randn(1, 2, D // 2, h // 2, w // 2, 8, device=dev).bfloat16())
    y = cr.conv11.forward_g8(x, skip=c0)
    impl = ops.conv_impl_for(8, 1, 1, False)
    t0 = timeit(lambda: cr.conv0.forward_g8(vol))
    t1 = timeit(lambda: cr.conv1.forward_g8(c0))
    t2 = timeit(lambda: cr.conv2.forward_g8(c1))
    t11 = timeit(lambda: cr.conv11.forward_g8(x, skip=c0))
    tp = timeit(lambda: ops.conv3d(y, cr._prob_prepared(impl), None, None, 1, 1, False, False, None, torch.float32, True, impl))
    print(f"stage{stage+1}: conv0 {t0:7.1f}  conv1 {t1:7.1f}  conv2 {t2:7.1f}  conv11 {t11:7.1f}  prob {tp:7.1f} us")
