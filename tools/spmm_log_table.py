#!/usr/bin/env python
"""Condense a tools/spmm_bench.py log (one JSON object per line) into a table: python tools/spmm_log_table.py LOG"""
import json
import sys

for ln in open(sys.argv[1]):
    if ln.startswith("=="):
        print(ln.strip())
    elif ln.startswith("{"):
        d = json.loads(ln)
        if "variant" in d:
            print("%-6s %-8s rows=%-7d C=%-4d %7.1f us (best %6.1f)  %6.0f GB/s  frac %.3f"
                  % (d["op"], d["variant"], d["rows"], d["C"], d["us"], d["us_best"], d["GBps"], d["frac_of_measured_peak"]))
    elif "Error" in ln or "error" in ln:
        print(ln.strip()[:200])
