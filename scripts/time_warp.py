"""Device time of the warp+aggregate kernel per stage at the BASELINE shapes."""
import sys
import torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
from damvsnet_b200 import synthetic
from damvsnet_b200.runner import HotPathRunner, make_workload
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
stages = make_workload(1152, 1600, 5, [48, 32, 8], seed=0, device=dev)
tot = 0
for s, (f, p, d) in enumerate(stages):
    nh = [dm.ops.features_to_nhwc(x) for x in f]
    rt = runner.depthnet.stage_rot_trans(p)
    wn = runner.depthnet.weight_net[s].folded()
    for _ in range(3):
        dm.ops.warp_aggregate(nh[0], nh[1:], rt, d, wn, "adaptive", torch.bfloat16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dm.ops.warp_aggregate(nh[0], nh[1:], rt, d, wn, "adaptive", torch.bfloat16)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot += ms
    print(f"stage{s+1}: {ms*1e3:.1f} us")
print(f"total {tot*1e3:.1f} us")
