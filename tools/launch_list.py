#!/usr/bin/env python
"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` launch list to (id, kernel, grid, block, ns) rows and a
per-kernel aggregate (launches, total / mean time, share of the listed time).

    python tools/launch_list.py RAW.csv OUT_PREFIX      ->  OUT_PREFIX.csv, OUT_PREFIX_summary.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)                 # drop the argument list
    name = name.replace("void ", "")
    return name if len(name) <= 90 else name[:87] + "..."


def main():
    raw, prefix = sys.argv[1], sys.argv[2]
    rows = []
    with open(raw) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((int(r["ID"]), short(r["Kernel Name"]), r["Grid Size"], r["Block Size"],
                     float(r["Metric Value"].replace(",", ""))))
    with open(prefix + ".csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["id", "kernel", "grid", "block", "ns"])
        w.writerows(rows)
    agg = OrderedDict()
    total = sum(r[4] for r in rows)
    for _, k, _, _, ns in rows:
        a = agg.setdefault(k, {"launches": 0, "total_us": 0.0})
        a["launches"] += 1
        a["total_us"] += ns / 1e3
    out = sorted(({"kernel": k, "launches": v["launches"], "total_us": round(v["total_us"], 1),
                   "mean_us": round(v["total_us"] / v["launches"], 2), "share": round(v["total_us"] * 1e3 / total, 4)}
                  for k, v in agg.items()), key=lambda d: -d["total_us"])
    json.dump({"launches": len(rows), "total_us": round(total / 1e3, 1), "kernels": out}, open(prefix + "_summary.json", "w"),
              indent=1)
    for d in out[:25]:
        print("%6d  %9.1f us  %5.1f %%  %s" % (d["launches"], d["total_us"], 100 * d["share"], d["kernel"]))


if __name__ == "__main__":
    main()
