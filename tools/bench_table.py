#!/usr/bin/env python
"""Print the per-kernel table of a bench.py JSON line (tools/bench_table.py gpurun_out/xxx.json)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("ms/step %.3f  eager %.3f  launches %s" % (d["ms_per_step"], d["ms_per_step_eager"], d["gpu_launches"]))
for k in ("e2e", "e2e_coo_upload", "e2e_gpu_built_operators"):
    if d.get(k):
        print("  %-26s %.3f ms  %.0f meshes/s  h2d %.1f MB" % (k, d[k]["ms_per_step"], d[k]["value"], d[k]["h2d_bytes_per_step"] / 1e6))
for k, v in d.get("rooflines", {}).items():
    print("  family %-24s frac %.3f  avg %.1f us x %d  = %.3f ms/step" % (k, v["frac"], v["avg_launch_us"], v["launches_timed"] // d["steps"], v["ms_per_step_eager"]))
tot = 0
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    f = v.get("frac_of_hbm_peak")
    print("%-78s %5.1f x %7.1f us = %6.3f ms  %s" % (k[:78], v["launches_per_step"], v["us_per_launch"], v["ms_per_step"], "" if f is None else "%.2f" % f))
print("sum of sn kernels (eager events): %.3f ms" % tot)
