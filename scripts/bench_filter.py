"""Fused geometric-consistency filter at the DTU-test size (1152x1600, 4 source views): device time, HBM fraction,
and the numpy oracle (the reference's CPU algorithm) timed beside it -- SURVEY.md 8f rank 2."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
from damvsnet_b200 import fusion
from oracle import damvs_oracle as O
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
t0 = time.perf_counter()
want = O.filter_reference_view(depths[0], confs, Ks[0], Es[0], depths[1:], Ks[1:], Es[1:])
cpu_s = time.perf_counter() - t0
mism = float((out["final_mask"].cpu().numpy() != want["final_mask"]).mean())
print(json.dumps({"component": "geometric-consistency filter, 1 reference view x 4 source views, 1152x1600", "gpu_ms": ms, "views_per_s": 1e3 / ms,
                  "algorithmic_GBps": nbytes / ms / 1e6, "cpu_oracle_s": cpu_s, "cpu_threads": 1, "final_mask_mismatch_vs_oracle": mism}))
