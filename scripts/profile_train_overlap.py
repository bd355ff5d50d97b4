#!/usr/bin/env python
"""Timeline of the training step's collectives (BASELINE.json configs[3]): where the per-stage NCCL
gradient all-reduces sit relative to the backward kernels.

nsys is not in this image; torch.profiler (Kineto / CUPTI) records every GPU kernel of one training step with device
timestamps.  The script reports, on rank 0: the step's GPU span, each all-reduce kernel's start / duration as an
offset into that span, and how much of each lies under kernels of this library (overlap).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 \
        scripts/profile_train_overlap.py [--out gpurun_out/train_overlap.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_overlap.json"))
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import damvsnet_b200 as dm
    from damvsnet_b200 import synthetic
    from damvsnet_b200.training import HotPathTrainer
    sys.path.insert(0, ROOT)
    import bench
    dm.set_precision("bf16")
    trainer = HotPathTrainer(synthetic.hot_path_state_dict(seed=0), device=dev)
    stages = bench.device_workload(512, 640, 5, [48, 32, 8], dev, seed=rank, batch=4)
    stages = [([f.requires_grad_(True) for f in fs], p, d) for fs, p, d in stages]
    gts = [(d[:, d.shape[1] // 2] + 0.5).contiguous() for _, _, d in stages]
    masks = [torch.ones_like(g) for g in gts]

    def step():
        for fs, _, _ in stages:
            for f in fs:
                f.grad = None
        trainer.train_step(stages, gts, masks)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
        ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in ev), key=lambda t: t[0])
        t0, t1 = ks[0][0], max(k[1] for k in ks)
        nccl = [k for k in ks if "nccl" in k[2].lower()]
        ours = [k for k in ks if "damvs" in k[2] or "tcf::" in k[2] or "tcp::" in k[2]]
        rows = []
        for s, e, name in nccl:
            under = 0.0
            for a, b, _ in ours:
                lo, hi = max(s, a), min(e, b)
                if hi > lo:
                    under += hi - lo
            # which backward kernels run while it is in flight
            rows.append({"kernel": name[:60], "start_ms": (s - t0) / 1e3, "duration_ms": (e - s) / 1e3,
                         "overlapped_by_library_kernels_ms": min(under, e - s) / 1e3})
        last_ours = max(b for _, b, _ in ours)
        out = {"world": world, "step_gpu_span_ms": (t1 - t0) / 1e3, "library_kernels": len(ours),
               "library_kernel_busy_ms": sum(b - a for a, b, _ in ours) / 1e3,
               "last_library_kernel_end_ms": (last_ours - t0) / 1e3, "allreduce": rows,
               "note": "offsets are into the GPU span of ONE training step (forward + backward + all-reduce + Adam) on rank 0; the "
                       "per-stage buckets are all-reduced from gradient hooks, stage 3 first, while the backward kernels of the "
                       "earlier stages still run"}
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
