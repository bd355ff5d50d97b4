import sys, time
import torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
from damvsnet_b200 import synthetic
from damvsnet_b200.runner import HotPathRunner, make_workload
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
host = make_workload(1152, 1600, 5, [48, 32, 8], seed=0)
pinned = runner.pin_stages(host)
for _ in range(3): runner.run_host(pinned)
torch.cuda.synchronize()
# instrument: wrap run_stage to record events
evs = []
orig = runner.run_stage
def rs(i, f, p, d):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); out = orig(i, f, p, d); b.record(); evs.append((i, a, b)); return out
runner.run_stage = rs
base = torch.cuda.Event(enable_timing=True); base.record()
t0 = time.perf_counter()
tks = []
for v in range(3):
    tks.append(runner.submit_host(pinned)); print(f"view {v} submitted at host t={1e3*(time.perf_counter()-t0):.2f} ms")
for t in tks:
    runner.collect(t); print(f"collected at host t={1e3*(time.perf_counter()-t0):.2f} ms")
torch.cuda.synchronize()
for i, a, b in evs:
    print(f"stage {i}: start {base.elapsed_time(a):.2f} end {base.elapsed_time(b):.2f} ms")
