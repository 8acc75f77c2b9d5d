#!/usr/bin/env python
"""Attribute executed SASS instructions of one kernel in an .ncu-rep to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled-name-substring> [units]

Uses `ncu --page source --csv` (per-SASS-instruction executed counts, needs --import-source / -lineinfo) and
`nvdisasm -g` on the cubin embedded in the object (line table), aligned by instruction order.
`units` divides the per-line counts (e.g. the number of blocks) to print instructions per unit of work.
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def main():
    rep, obj, name = sys.argv[1:4]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = [r for r in rows if r and r[0] == "Address"][0]
    data = [r for r in rows if r and r[0].startswith("0x")]
    ie, samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    sass = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(os.path.join(tmp, "*.cubin")), capture_output=True,
                          text=True).stdout.splitlines()
    seq, cur, on, files = [], None, False, {}
    for ln in sass:
        if ln.startswith("//--------------------- .text."):
            on = name in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            files[cur[0]] = m.group(1)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            seq.append(cur)
    if len(seq) != len(data):
        print("warning: %d SASS instructions in the object vs %d in the report (stale build?)" % (len(seq), len(data)))
    by, bs = Counter(), Counter()
    for key, r in zip(seq, data):
        by[key] += int(r[ie])
        bs[key] += int(r[samp])
    tot, tots = sum(by.values()), max(sum(bs.values()), 1)
    src = {f: open(p).read().split("\n") for f, p in files.items() if os.path.exists(p)}
    print("total executed %d (%.1f per unit), samples %d" % (tot, tot / units, tots))
    for key, n in by.most_common(28):
        f, line = key if key else ("?", 0)
        text = src.get(f, [""] * (line + 1))[line - 1].strip() if line else "?"
        print("%-22s %9d %5.1f%% %6.2f/unit  stall %4.1f%%  %s" % ("%s:%d" % (f, line), n, 100.0 * n / tot, n / units,
                                                              100.0 * bs[key] / tots, text[:90]))


if __name__ == "__main__":
    main()
