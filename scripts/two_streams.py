"""Throughput with k views in flight on k streams (each a captured 3-stage graph) -- experiment."""
import sys, torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
from damvsnet_b200 import synthetic
from damvsnet_b200.runner import HotPathRunner, make_workload
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sd = synthetic.hot_path_state_dict(seed=0)
runners, stages, streams, graphs = [], [], [], []
for k in range(K):
    r = HotPathRunner(sd, device=dev)
    st = make_workload(1152, 1600, 5, [48, 32, 8], seed=k, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        r.run_device(st); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            outs = r.run_device(st)
    runners.append(r); stages.append(st); streams.append(s); graphs.append((g, outs))
torch.cuda.synchronize()
def run(n):
    for i in range(n):
        k = i % K
        with torch.cuda.stream(streams[k]):
            graphs[k][0].replay()
run(2 * K); torch.cuda.synchronize()
N = 40
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in streams: s.wait_event(a)
run(N)
for s in streams:
    e = torch.cuda.Event(); e.record(s); torch.cuda.current_stream().wait_event(e)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(f"K={K}: {ms / N:.3f} ms/view, {N / ms * 1e3:.1f} views/s")
