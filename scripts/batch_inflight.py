"""Throughput for (batch B per launch) x (K graphs in flight) -- experiment."""
import sys, torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import synthetic
from damvsnet_b200.runner import HotPathRunner, ViewPipeline, make_workload
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
sd = synthetic.hot_path_state_dict(seed=0)
for B, K in ((1, 1), (1, 4), (2, 1), (2, 2), (2, 3), (4, 1), (4, 2)):
    runner = HotPathRunner(sd, device=dev)
    sets = [make_workload(1152, 1600, 5, [48, 32, 8], batch=B, seed=k, device=dev) for k in range(K)]
    pipe = ViewPipeline(runner, sets)
    pipe.fork(); pipe.submit(2 * K); pipe.join(); torch.cuda.synchronize()
    n = 24 // B
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pipe.fork(); pipe.submit(n); pipe.join(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / (n * B)
    print(f"B={B} K={K}: {ms:.3f} ms/view  {1e3 / ms:.1f} views/s", flush=True)
    del pipe, sets, runner
    torch.cuda.empty_cache()
