#!/usr/bin/env python
"""Digest `ncu --page raw --csv` files into the per-kernel figures bench.py and the docs cite: duration, DRAM bytes (read + write)
per launch, DRAM / tensor-pipe / SM / XU utilisation.  usage: ncu_digest.py out.json name=file.csv ..."""
import csv
import json
import sys

WANT = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
        "lts__t_bytes.sum": "l2_bytes", "launch__grid_size": "grid", "launch__registers_per_thread": "regs",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct"}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def digest(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")][:60]}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] not in ("", "n/a"):
                d[WANT[h]] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        out.append(d)
    return out


res = {}
for arg in sys.argv[2:]:
    name, path = arg.split("=")
    ls = digest(path)
    res[name] = {"launches_captured": len(ls), "duration_us": [round(x.get("duration", 0), 2) for x in ls],
                 "dram_bytes_per_launch": sum(x.get("dram_read", 0) + x.get("dram_write", 0) for x in ls) / max(len(ls), 1),
                 "dram_pct": [round(x.get("dram_pct", 0), 2) for x in ls], "tensor_pipe_pct": [round(x.get("tensor_pipe_pct", 0), 2) for x in ls], "issue_pct": [round(x.get("issue_pct", 0), 2) for x in ls],
                 "sm_pct": [round(x.get("sm_pct", 0), 2) for x in ls], "xu_pct": [round(x.get("xu_pct", 0), 2) for x in ls],
                 "l2_bytes_per_launch": sum(x.get("l2_bytes", 0) for x in ls) / max(len(ls), 1), "grid": [int(x.get("grid", 0)) for x in ls],
                 "regs": [int(x.get("regs", 0)) for x in ls], "kernel": ls[0]["kernel"] if ls else ""}
json.dump(res, open(sys.argv[1], "w"), indent=1)
print(json.dumps(res, indent=1))
