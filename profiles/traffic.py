#!/usr/bin/env python
"""ncu launch list with DRAM counters -> profiles/ncu_traffic.json (what bench.py reports as `roofline.traffic`).

Capture (one B200, after the same command exited 0 without ncu), all launches of ONE training step after 3 warm-ups:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/r2_traffic_cfg2.csv python bench.py --steps 1 --warmup 3 --profile-steps 0 --no-cpu-baseline

    python profiles/traffic.py cfg2 gpurun_out/r2_traffic_cfg2.csv <edges per step> <message-passing steps> <launches per training step>

Only the LAST training step of the capture is used (the tail of the launch list: `launches per training step` kernels of
this library, counted by bench.py's gpu_launches).  Bytes are summed per kernel name and divided by edges x
message-passing steps."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ncu_traffic.json")


def main():
    workload, path, edges, mp_steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(r for r in rows if "Kernel Name" in r)
    body = rows[rows.index(hdr) + 1:]
    ik, im, iv, iu, iid = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per_launch = collections.OrderedDict()
    for r in body:
        if len(r) <= iv:
            continue
        d = per_launch.setdefault(r[iid], {"name": r[ik]})
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
        d[r[im]] = v * scale
    launches = list(per_launch.values())
    ours = [l for l in launches if "at::" not in l["name"] and "nccl" not in l["name"].lower()]
    n_step = int(sys.argv[5]) if len(sys.argv) > 5 else len(ours)
    step = ours[-n_step:]
    agg = collections.OrderedDict()
    for l in step:
        name = re.sub(r"\(.*", "", l["name"])
        name = re.sub(r"<unnamed>::|void ", "", name)
        a = agg.setdefault(name, {"launches": 0, "bytes": 0.0, "us": 0.0})
        a["launches"] += 1
        a["bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        a["us"] += l.get("gpu__time_duration.sum", 0.0)
    tree = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=HERE).stdout.strip()
    doc = json.load(open(OUT)) if os.path.exists(OUT) else {"workloads": {}}
    doc["tree"] = tree
    doc["workloads"][workload] = {
        "capture": os.path.basename(path), "edges_per_step": edges, "message_passing_steps": mp_steps,
        "launches_in_step": len(step),
        "kernels": {k: {"launches_per_step": v["launches"] / mp_steps,
                        "dram_bytes_per_edge_step": v["bytes"] / (edges * mp_steps),
                        "us_per_training_step_under_ncu": v["us"]} for k, v in agg.items()}}
    json.dump(doc, open(OUT, "w"), indent=1)
    tot = sum(v["bytes"] for v in agg.values())
    print(f"{workload}: {len(step)} launches, {tot / 1e9:.2f} GB DRAM traffic in the step -> {OUT}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["bytes"])[:12]:
        print(f"  {k[:60]:60s} x{v['launches']:3d}  {v['bytes'] / (edges * mp_steps):9.1f} B/edge-step  {v['us'] / 1e3:8.2f} ms")


if __name__ == "__main__":
    main()
