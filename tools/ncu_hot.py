#!/usr/bin/env python
"""Top stall-sample instructions of a kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass` output."""
import csv
import sys


def main(path, kernel_idx=0, top=40):
    rows = list(csv.reader(open(path)))
    # split into kernels
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[kernel_idx]
    e = starts[kernel_idx + 1] if kernel_idx + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    body = rows[s + 2:e]
    ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[isamp] or 0) for r in body)
    print(rows[s][1][:100], "total samples", total)
    base = int(body[0][ia], 16)
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:top]
    for i in sorted(ranked):
        r = body[i]
        st = sorted(((int(r[c] or 0), h[6:]) for c, h in stall_cols), reverse=True)[:3]
        print("%5x %6d %5.1f%% exec %8s  %-60s %s" % (int(r[ia], 16) - base, int(r[isamp]), 100.0 * int(r[isamp]) / max(total, 1), r[iexec],
                                               r[isrc].strip()[:60], " ".join("%s:%d" % (h, n) for n, h in st if n)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
