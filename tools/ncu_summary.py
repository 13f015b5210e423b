#!/usr/bin/env python
"""Condense ncu output into the small text summaries kept under profiles/.

    python tools/ncu_summary.py report  <file.ncu-rep> [...]   # one block per profiled kernel launch
    python tools/ncu_summary.py launches <launches.csv>        # per-kernel count / total / share from the
                                                               # `--metrics gpu__time_duration.sum` launch list
Reads reports with `ncu -i ... --page raw --csv` (ncu is in the image; no GPU needed to read).
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def raw_rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[start], rows[start + 1], rows[start + 2:]


def report(path):
    hdr, units, launches = raw_rows(path)
    blocks = []
    for vals in launches:
        if len(vals) != len(hdr):
            continue
        d = dict(zip(hdr, zip(units, vals)))
        blk = OrderedDict()
        blk["report"] = path.split("/")[-1]
        blk["kernel"] = d.get("Kernel Name", ("", "?"))[1][:120]
        for k in KEEP:
            hit = [h for h in hdr if h == k or h.endswith("." + k)]
            if hit and d[hit[0]][1] != "":
                u, v = d[hit[0]]
                blk[k] = f"{v} {u}".strip()
        blocks.append(blk)
    return blocks


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        name = r[ki].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        name = (name.split("<")[0] + ("<" + name.split("<", 1)[1][:12] if "<" in name and name.split("<")[0].startswith("tc_") else "")).strip()[:60]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [{"kernel": k, "launches": a[0], "total_us": round(a[1], 1), "avg_us": round(a[1] / a[0], 2),
            "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    return {"file": path.split("/")[-1], "total_us": round(tot, 1), "kernels": out}


if __name__ == "__main__":
    mode, paths = sys.argv[1], sys.argv[2:]
    if mode == "report":
        for p in paths:
            for b in report(p):
                print(json.dumps(b, indent=1))
    else:
        for p in paths:
            print(json.dumps(launches(p), indent=1))
