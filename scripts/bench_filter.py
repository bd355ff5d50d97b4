"""Fused geometric-consistency filter at the DTU-test size (1152x1600, 4 source views): device time and algorithmic
bandwidth -- SURVEY.md 8f rank 2.  (The numpy oracle of the reference's CPU algorithm is timed beside it by
tests/test_gpu_filter.py::test_full_size_against_oracle; this script stays clear of oracle/.)"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from damvsnet_b200 import fusion
from make_golden_filter import make_scene
H, W, n = 1152, 1600, 4
Ks, Es, depths = make_scene(3, H, W, n)
rs = np.random.RandomState(0)
confs = [rs.rand(H, W).astype(np.float32) for _ in range(3)]
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(a).to(dev)
td, tc = [t(x) for x in depths], [t(x) for x in confs]
fn = lambda: fusion.filter_reference_view(td[0], tc, Ks[0], Es[0], td[1:], Ks[1:], Es[1:])
out = fn(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): fn()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
nbytes = (1 + 3 + n) * H * W * 4 + H * W * (4 + 3)
print(json.dumps({"component": "geometric-consistency filter, 1 reference view x 4 source views, 1152x1600", "gpu_ms": ms, "views_per_s": 1e3 / ms,
                  "algorithmic_GBps": nbytes / ms / 1e6, "final_mask_fraction": float(out["final_mask"].float().mean())}))
