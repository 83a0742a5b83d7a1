#!/usr/bin/env python3
"""Hot source lines of one kernel of an ncu report captured with --import-source on (-lineinfo builds):
  python profiles/ncu_lines.py report.ncu-rep kernel_regex [min_pct]
Prints, per CUDA source line, its share of executed warp instructions and of stall samples, and the top stall reasons."""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows, hdr, fname = [], None, ""
for line in csv.reader(io.StringIO(txt)):
    if len(line) == 2 and line[0] == "File Path":
        fname = line[1].split("/")[-1]
    elif len(line) > 5 and line[0] == "Line No":
        hdr = line
    elif hdr and len(line) == len(hdr) and line[2] == "-":
        d = dict(zip(hdr[4:], line[4:]))
        d["line"], d["src"], d["file"] = line[0], line[1], fname
        rows.append(d)


def I(r, k):
    try:
        return int(r[k])
    except Exception:
        return 0


tot = sum(I(r, "Instructions Executed") for r in rows) or 1
ts = sum(I(r, "# Samples") for r in rows) or 1
print("warp instructions %d, samples %d" % (tot, ts))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows:
    n, s = I(r, "Instructions Executed"), I(r, "# Samples")
    if 100.0 * n / tot >= min_pct or 100.0 * s / ts >= min_pct:
        top = sorted(((I(r, k), k[6:]) for k in stalls), reverse=True)[:3]
        print("%-16s:%-4s inst %5.1f%% samp %5.1f%%  %-34s | %s" % (r["file"], r["line"], 100.0 * n / tot, 100.0 * s / ts,
              " ".join("%s=%d" % (k, v) for v, k in top if v), r["src"].strip()[:100]))
