#!/usr/bin/env python
"""One table row per kernel launch from an `ncu --set full` report (the numbers quoted in DESIGN.md section 3).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/summarize_ncu_full.py /tmp/raw.csv > profiles/rNN_ncu_summary.md
"""
import csv
import sys

COLS = (  # (column title, candidate metric names, scale, format)
    ("us", ("gpu__time_duration.sum",), 1.0, "{:.1f}"),
    ("rd MB", ("dram__bytes_read.sum",), 1.0, "{:.1f}"),
    ("wr MB", ("dram__bytes_write.sum",), 1.0, "{:.1f}"),
    ("dram %", ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 1.0, "{:.1f}"),
    ("sm %", ("sm__throughput.avg.pct_of_peak_sustained_elapsed",), 1.0, "{:.1f}"),
    ("l1tex %", ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), 1.0, "{:.1f}"),
    ("tensor pipe %", ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
                       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), 1.0, "{:.1f}"),
    ("issue %", ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active"), 1.0, "{:.1f}"),
    ("warps active %", ("sm__warps_active.avg.pct_of_peak_sustained_active",), 1.0, "{:.1f}"),
    ("regs", ("launch__registers_per_thread",), 1.0, "{:.0f}"),
    ("warp inst", ("smsp__inst_executed.sum",), 1.0, "{:.0f}"),
)
UNIT_SCALE = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    print("| kernel | grid | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:48]
        grid = r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""
        cells = []
        for _, cands, _, fmt in COLS:
            cell = ""
            for m in cands:
                if m in hdr:
                    try:
                        v = float(r[hdr.index(m)].replace(",", ""))
                    except ValueError:   # "n/a", "no data"
                        continue
                    v *= UNIT_SCALE.get(units[hdr.index(m)], 1.0)
                    cell = fmt.format(v)
                    break
            cells.append(cell)
        print(f"| `{name}` | {grid} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
