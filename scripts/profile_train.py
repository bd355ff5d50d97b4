"""Kernel-level breakdown of one training step (torch.profiler, CUDA activities) -- development aid."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import damvsnet_b200 as dm
dm.set_precision("bf16")   # measures the reduced-precision pipeline (the package default is fp32)
from damvsnet_b200 import synthetic
from damvsnet_b200.runner import make_workload
from damvsnet_b200.training import HotPathTrainer

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 4))
trainer = HotPathTrainer(synthetic.hot_path_state_dict(seed=0), device=dev)
stages = make_workload(512, 640, 5, [48, 32, 8], batch=B, seed=0, device=dev)
stages = [([f.requires_grad_(True) for f in fs], p, d) for fs, p, d in stages]
gts = [d[:, d.shape[1] // 2].contiguous() + 0.5 for _, _, d in stages]
masks = [torch.ones_like(g) for g in gts]
for _ in range(2):
    trainer.train_step(stages, gts, masks)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    trainer.train_step(stages, gts, masks)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
