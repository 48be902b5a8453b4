#!/usr/bin/env python
"""Turns gpurun_out ncu artefacts into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/X_launches.csv > profiles/X_launches_summary.csv
  python profiles/summarize.py raw gpurun_out/X_prof.ncu-rep     > profiles/X_prof_raw.csv
"""
import collections
import csv
import re
import subprocess
import sys

RAW = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        name = re.sub(r".*::", "", name)
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + v)
    tot = sum(t for _, t in agg.values())
    print("kernel,launches,total_us,share_of_captured_window")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"\"{k}\",{n},{t:.1f},{t / tot:.4f}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = [(w, hdr.index(w)) for w in RAW if w in hdr]
    print(",".join(w for w, _ in idx))
    print(",".join(rows[1][i] for _, i in idx))
    for r in rows[2:]:
        print(",".join('"' + re.sub(r"\(.*", "", r[i]) + '"' if w == "Kernel Name" else r[i] for w, i in idx))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
