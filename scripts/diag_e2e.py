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
nbytes = runner.h2d_bytes(host)
def sync(): torch.cuda.synchronize()
# (a) H2D only
bufs = [([torch.empty_like(f, device=dev) for f in fs], torch.empty_like(p, device=dev), torch.empty_like(d, device=dev)) for fs, p, d in pinned]
for rep in range(3):
    sync(); t0 = time.perf_counter()
    for (fs, p, d), (bf, bp, bd) in zip(pinned, bufs):
        for a, b in zip(fs, bf): b.copy_(a, non_blocking=True)
        bp.copy_(p, non_blocking=True); bd.copy_(d, non_blocking=True)
    sync(); t = time.perf_counter() - t0
print(f"H2D only: {t*1e3:.2f} ms  {nbytes/t/1e9:.1f} GB/s")
# (b) compute only
dstages = bufs
for _ in range(3): runner.run_device(dstages)
sync(); t0 = time.perf_counter()
for _ in range(5): runner.run_device(dstages)
sync(); print(f"compute only (eager): {(time.perf_counter()-t0)/5*1e3:.2f} ms")
# host launch cost of one run_device (no sync)
t0 = time.perf_counter(); runner.run_device(dstages); t1 = time.perf_counter(); sync()
print(f"host time to enqueue one view: {(t1-t0)*1e3:.2f} ms")
# (c) submit/collect
for depth in (1, 2):
    for _ in range(2): runner.run_host(pinned)
    sync(); t0 = time.perf_counter()
    pend = []
    n = 8
    for i in range(n):
        pend.append(runner.submit_host(pinned))
        if len(pend) >= depth:
            t = pend.pop(0); runner.collect(t); runner.release(t)
    for t in pend: runner.collect(t); runner.release(t)
    sync(); print(f"pipeline depth {depth}: {(time.perf_counter()-t0)/n*1e3:.2f} ms/view; device allocs {torch.cuda.memory_stats()['num_device_alloc']}, reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB")
