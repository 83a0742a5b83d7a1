#!/usr/bin/env python3
"""SASS instructions (with executed counts and stall samples) of the source lines [lo, hi] of one kernel:
  python profiles/ncu_sass.py report.ncu-rep kernel_regex file_suffix lo hi"""
import csv, io, subprocess, sys
rep, kern, fsuf, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
hdr, fname, cur = None, "", None
for line in csv.reader(io.StringIO(txt)):
    if len(line) == 2 and line[0] == "File Path":
        fname = line[1]
    elif len(line) > 5 and line[0] == "Line No":
        hdr = line
    elif hdr and len(line) == len(hdr) and fname.endswith(fsuf):
        try:
            ln = int(line[0])
        except ValueError:
            continue
        if lo <= ln <= hi:
            d = dict(zip(hdr[4:], line[4:]))
            if line[2] == "-":
                print("---- %d: %s" % (ln, line[1].strip()[:110]))
            else:
                st = sorted(((int(d[k]) if d[k].isdigit() else 0, k[6:]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:3]
                print("   %-60s exec %10s samp %7s  %s" % (line[3].strip()[:60], d["Instructions Executed"], d["# Samples"],
                      " ".join("%s=%d" % (k, v) for v, k in st if v)))
