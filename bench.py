#!/usr/bin/env python
"""Benchmark of the DA-MVSNet cost-volume hot path (BASELINE.json metric: views/sec).

A "step" is one reference view's three-stage hot path (fused warp+aggregate ->
CostRegNet -> softmax/regression head, models/cas_mvsnet.py:18-134 of the
reference, three times) on synthetic DTU-test-shaped inputs: 1152x1600, N=5,
D=48/32/8, batch 1 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]        # this implementation
    python bench.py --impl reference ...                         # CPU oracle port on the host cores

One process per GPU (torchrun for N>1); reference views are independent, so ranks
share nothing on the data path and the scaling is weak (K views per rank).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "views/sec at 1152x1600 N=5 D=48/32/8 (hot path: warp+aggregate, CostRegNet, regression head)"
UNIT = "views/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--conv-impl", default="auto", choices=["auto", "direct", "tcgen05"])
    ap.add_argument("--mode", default="adaptive", choices=["adaptive", "variance"])
    ap.add_argument("--height", type=int, default=1152)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--nviews", type=int, default=5)
    ap.add_argument("--ndepths", default="48,32,8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--detail", action="store_true", help="print a per-layer device-time table to stderr")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--inflight", type=int, default=4, help="independent views in flight on separate streams (graph mode)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor": p["bf16_tflops"], "tensor_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference).  The only place bench.py touches oracle/.
# ---------------------------------------------------------------------------------------------
def cpu_oracle_step_fn(args, crop_h, crop_w):
    from damvsnet_b200 import synthetic
    from damvsnet_b200.runner import make_workload
    from oracle import damvs_oracle as O
    nd = [int(x) for x in args.ndepths.split(",")]
    sd = synthetic.hot_path_state_dict(seed=0, mode=args.mode)
    stages = make_workload(crop_h, crop_w, args.nviews, nd, seed=0)

    def step():
        with torch.no_grad():
            for s, (f, p, d) in enumerate(stages):
                O.depthnet_forward(s, f, p, d, sd, args.mode)
    return step


def cpu_leg(args, steps, warmup, budget_s):
    """Time the CPU oracle port with all host threads on a bounded crop of the workload.
    views/s = (crop area / full area) / seconds per crop."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = args.height * args.width
    probe = cpu_oracle_step_fn(args, 128, 160)
    probe()
    t0 = time.perf_counter()
    probe()
    t_probe = time.perf_counter() - t0
    choice = (128, 160)
    for ch, cw in ((384, 512), (256, 320), (128, 160)):
        if ch > args.height or cw > args.width:
            continue
        est = t_probe * (ch * cw) / (128 * 160) * (steps + warmup)
        if est <= budget_s:
            choice = (ch, cw)
            break
    step = cpu_oracle_step_fn(args, *choice)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    ratio = choice[0] * choice[1] / full
    return {"value": ratio / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{choice[0]}x{choice[1]} crop of the {args.height}x{args.width} view (all 3 stages, N={args.nviews}, "
                      f"D={args.ndepths}), {steps} timed passes, scaled by area ratio {ratio:.4f}; torch-CPU fp32 oracle port, "
                      f"{cores} threads",
            "sec_per_sample": sec}


def config_of(args, extra=None):
    c = {"workload": f"DTU-test {args.height}x{args.width}, N={args.nviews}, D={args.ndepths.replace(',', '/')}, batch 1, "
                     f"agg={args.mode} (BASELINE.json configs[1])",
         "precision": args.precision,
         "l2": "no flush needed: per-step inputs (0.6 GB) and intermediates (>2 GB) exceed the 126 MB L2"}
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    r = cpu_leg(args, steps, warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args, {"precision": "fp32", "note": "CPU oracle port of the reference path (the reference is "
                                                                    "Python and /root/reference does not exist on the GPU box)"}),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import damvsnet_b200 as dm
    from damvsnet_b200 import _lib, synthetic
    from damvsnet_b200.runner import HotPathRunner, ViewPipeline, make_workload
    _lib.check(_lib.load().damvs_check_device(local))
    dm.set_precision(args.precision, args.conv_impl)
    nd = [int(x) for x in args.ndepths.split(",")]
    sd = synthetic.hot_path_state_dict(seed=0, mode=args.mode)
    runner = HotPathRunner(sd, mode=args.mode, device=dev)
    host_stages = make_workload(args.height, args.width, args.nviews, nd, seed=rank)
    dev_stages = [([f.to(dev) for f in fs], p.to(dev), d.to(dev)) for fs, p, d in host_stages]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    inflight = 1 if args.no_graph else max(1, args.inflight)
    pipe = None
    if not args.no_graph:
        # K independent views in flight: K input sets (distinct data), one captured graph + stream each
        # (further sets are made on the device from the first one: features shifted by a few pixels + a small offset, so
        # the data differ per view without another minute of host-side synthesis)
        sets = [dev_stages]
        for k in range(1, inflight):
            sets.append([([torch.roll(f, shifts=(3 * k, 5 * k), dims=(2, 3)).add_(0.01 * k).contiguous() for f in fs], p.clone(),
                          (d + 0.05 * k).contiguous()) for fs, p, d in dev_stages])
        pipe = ViewPipeline(runner, sets)

    def run_views(n):
        if pipe is None:
            for _ in range(n):
                runner.run_device(dev_stages)
        else:
            pipe.fork()
            pipe.submit(n)
            pipe.join()

    run_views(max(warmup, inflight))
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_views(steps)
    e1.record()
    barrier()
    launches = (_lib.launch_count() - l0) // steps
    if not args.no_graph:
        # graph replays do not pass through the C ABI: count the kernels the captured step contains
        n0 = _lib.launch_count()
        runner.run_device(dev_stages)
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    from damvsnet_b200 import sharding
    total_views, total_ms = sharding.reduce_throughput(steps, e0.elapsed_time(e1), dev)   # sum of views, max of device time
    value = total_views / (total_ms / 1e3)

    # ---- end to end: host buffers in, host results out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        pinned = runner.pin_stages(host_stages)
        warm = [runner.submit_host(pinned) for _ in range(3)]      # also allocates the pinned result buffers
        for t in warm:
            runner.collect(t)
            runner.release(t)
        e_steps = max(3, min(steps, 10))
        barrier()
        e0.record()
        # two views in flight: the upload of view i+1 overlaps the kernels of view i; every view's results are
        # read back to pinned host memory and waited for inside the timed region
        pending = None
        for _ in range(e_steps):
            t = runner.submit_host(pinned)
            if pending is not None:
                runner.collect(pending)
                runner.release(pending)
            pending = t
        runner.collect(pending)
        runner.release(pending)
        e1.record()
        barrier()
        e_views, e_ms = sharding.reduce_throughput(e_steps, e0.elapsed_time(e1), dev)
        e2e = {"value": e_views / (e_ms / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": runner.h2d_bytes(host_stages), "d2h_bytes_per_step": runner.d2h_bytes(host_stages),
               "steps": e_steps, "api": "HotPathRunner.submit_host/collect, 2 views in flight (pinned host features/proj/hypotheses in; "
                                        "depth, confidence, variance of 3 stages out to pinned host memory)"}
        del pinned

    # ---- roofline: per-call device times of the hot kernels, measured live with CUDA events
    roof, kernels, rooflines = None, None, None
    if rank == 0:
        pk = peaks()
        with dm.ops.CallTimer() as timer:
            for _ in range(3):
                runner.run_device(dev_stages)
        summ = timer.summary()
        if args.detail:
            for tag, d in timer.summary(detail=True).items():
                ms = d["ms"] / d["calls"]
                print(f"  {tag:48s} {ms * 1e3:9.1f} us  {d['bytes'] / d['calls'] / ms / 1e6:8.0f} GB/s  {d['flops'] / d['calls'] / ms / 1e9:8.1f} TF/s",
                      file=sys.stderr)
        kernels = {}
        for tag, d in summ.items():
            per_step_ms = d["ms"] / 3
            kernels[tag] = {"ms_per_step": per_step_ms, "launches_per_step": d["calls"] // 3,
                            "GBps": d["bytes"] / 3 / per_step_ms / 1e6, "TFLOPs": d["flops"] / 3 / per_step_ms / 1e9}
        # measured DRAM traffic per launch of each kernel class, from the committed ncu launch list
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            traffic = {k2: v["dram_bytes_per_launch"] for k2, v in json.load(open(tpath))["kernels"].items()}

        def roof_of(tag):
            k = kernels[tag]
            avg_ms = k["ms_per_step"] / k["launches_per_step"]
            if tag.startswith("conv3d"):
                # arithmetic intensity of the layer-by-layer CostRegNet is ~100 flop/B, below the ridge
                # (sustained bf16 / HBM = 207 flop/B): the HBM roof is the binding one; the tensor roof is reported beside it
                ai = (k["TFLOPs"] * 1e12) / max(k["GBps"] * 1e9, 1.0)
                return {"kernel": tag, "bound": "hbm", "achieved": k["GBps"], "peak": pk["hbm"], "unit": "GB/s",
                        "frac": k["GBps"] / pk["hbm"], "traffic": traffic.get(tag), "peak_source": pk["src"], "avg_launch_ms": avg_ms,
                        "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": pk["tensor_sustained"] * 1e3 / pk["hbm"],
                        "also_tensor": {"achieved": k["TFLOPs"], "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                                        "frac": k["TFLOPs"] / pk["tensor_sustained"], "peak_source": pk["src"] + " (sustained bf16)"}}
            return {"kernel": tag, "bound": "hbm", "achieved": k["GBps"], "peak": pk["hbm"], "unit": "GB/s",
                    "frac": k["GBps"] / pk["hbm"], "traffic": traffic.get(tag), "peak_source": pk["src"], "avg_launch_ms": avg_ms}

        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        roof = roof_of(top)
        rooflines = [roof_of(tag) for tag in kernels]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_leg(args, steps=2, warmup=1, budget_s=25.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic", "config": config_of(args, {"parallelism": f"views sharded x{world}, no data-path collective",
                                                                          "launch": "eager" if args.no_graph else f"cuda-graph replay of the 3-stage step, {inflight} independent views in flight on {inflight} streams"}),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "rooflines": rooflines if rank == 0 else None, "kernels": kernels,
                "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
