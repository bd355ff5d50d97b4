"""Device time of the warp+aggregate kernel per stage at the DTU-test shape (graph-captured) -- development aid."""
import sys, torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import ops, synthetic
from damvsnet_b200.runner import HotPathRunner, make_workload
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
stages = make_workload(1152, 1600, 5, [48, 32, 8], seed=0, device=dev)
out = []
for s, (f, p, d) in enumerate(stages):
    net = runner.depthnet
    rt = net.stage_rot_trans(p)
    nhwc = [ops.features_to_nhwc_half(x) for x in f]
    wnet = net.weight_net[s].folded()
    fn = lambda: ops.warp_aggregate(nhwc[0], nhwc[1:], rt, d, wnet, "adaptive", torch.bfloat16)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4): g.replay()
    b.record(); torch.cuda.synchronize()
    out.append(a.elapsed_time(b) / 20 * 1e3)
    rp = lambda: [ops.features_to_nhwc_half(x) for x in f]
    rp(); torch.cuda.synchronize()
    a.record()
    for _ in range(10): rp()
    b.record(); torch.cuda.synchronize()
    out.append(a.elapsed_time(b) / 10 * 1e3)
print("warp s1 %.0f (repack %.0f)  s2 %.0f (repack %.0f)  s3 %.0f (repack %.0f) us   total warp %.0f repack %.0f" % (*out, out[0] + out[2] + out[4], out[1] + out[3] + out[5]))
