#!/usr/bin/env python
"""Turn an ncu launch list (``ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv``
over one eager step of bench.py) into the per-kernel table profiles/rNN_launches_final.md and the per-class DRAM
traffic file profiles/rNN_traffic.json that bench.py's `roofline.traffic` reads.

    python scripts/summarize_launches.py gpurun_out/launches.csv profiles/r01 [launches per step] [last]

With `last`, the LAST `launches per step` launches of the hot-path classes are taken (one complete eager step at the end
of the capture, after set-up launches such as the hypothesis sampler and weight packing).
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict

CLASSES = (("nchw_to_nhwc", "repack"), ("warp_agg", "warp_agg"), ("conv3d_t", "conv"), ("head_", "head"), ("uncertainty", "hypotheses"))


def klass(name):
    for pat, k in CLASSES:
        if pat in name:
            return k
    return None


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    iid, iname, igrid, imet, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value"))
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[iid], {"name": r[iname], "grid": r[igrid]})
        d[r[imet]] = float(r[ival].replace(",", ""))
    ours = [(k, v) for k, v in launches.items() if klass(v["name"])]
    if len(sys.argv) > 4 and sys.argv[4] == "last":
        ours = [(k, v) for k, v in ours if klass(v["name"]) != "hypotheses"][-int(sys.argv[3]):]
    elif len(sys.argv) > 3:        # launches of this library in ONE step (the list may hold warm-up / further steps)
        ours = ours[:int(sys.argv[3])]
    keep = {k for k, _ in ours}
    if src != prefix + "_launches_final.csv":   # the committed copy holds only this library's launches of that step
        with open(prefix + "_launches_final.csv", "w", newline="") as f:
            w = csv.writer(f, quoting=csv.QUOTE_ALL)
            w.writerow(hdr)
            w.writerows(r for r in rows[1:] if r[iid] in keep)
    tot_ms = sum(v.get("gpu__time_duration.sum", 0.0) for _, v in ours) / 1e6
    agg = defaultdict(lambda: {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
    with open(prefix + "_launches_final.md", "w") as f:
        f.write("# One step of bench.py (eager), every launch of this library: ncu device time and DRAM bytes\n\n")
        f.write(f"From `{prefix}_launches_final.csv`. Per-launch times are cold-cache and serialised (ncu), so compare shares, not absolutes.\n\n")
        f.write("| # | kernel class | kernel | grid | ms | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|---|\n")
        for i, (k, v) in enumerate(ours):
            c = klass(v["name"])
            ms = v.get("gpu__time_duration.sum", 0.0) / 1e6
            rd, wr = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0)
            a = agg[c]
            a["n"] += 1; a["ms"] += ms; a["rd"] += rd; a["wr"] += wr
            short = re.sub(r"\s+", " ", v["name"])[:70]
            f.write(f"| {i} | {c} | `{short}` | {v['grid']} | {ms:.4f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} |\n")
        f.write("\n## Shares\n\n| kernel class | launches | ms (ncu) | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
        for c, a in agg.items():
            f.write(f"| {c} | {a['n']} | {a['ms']:.4f} | {a['ms'] / tot_ms:.3f} | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} |\n")
    out = {"source": f"{prefix}_launches_final.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none over `bench.py --no-graph`; one eager step; cold-cache serialised launches: compare shares)",
           "kernels": {c: {"launches_per_step": a["n"], "ms_per_step_ncu": round(a["ms"], 4), "share_of_step": round(a["ms"] / tot_ms, 4),
                           "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"], "dram_read_per_step": a["rd"], "dram_write_per_step": a["wr"]}
                       for c, a in agg.items()}}
    json.dump(out, open(prefix + "_traffic.json", "w"), indent=1)
    print(json.dumps(out["kernels"], indent=1))


if __name__ == "__main__":
    main()
