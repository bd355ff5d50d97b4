import sys, time
import torch
sys.path.insert(0, ".")
import damvsnet_b200 as dm
from damvsnet_b200 import synthetic
from damvsnet_b200 import runner as R
from damvsnet_b200.runner import HotPathRunner, make_workload
dev = torch.device("cuda:0")
runner = HotPathRunner(synthetic.hot_path_state_dict(seed=0), device=dev)
host = make_workload(1152, 1600, 5, [48, 32, 8], seed=0)
pinned = runner.pin_stages(host)
for _ in range(3): runner.run_host(pinned)
torch.cuda.synchronize()
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
t = runner.submit_host(pinned)
t2 = runner.submit_host(pinned)
pr.disable()
runner.collect(t); runner.collect(t2)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
