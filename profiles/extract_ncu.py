#!/usr/bin/env python3
"""Turns ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python profiles/extract_ncu.py launches gpurun_out/launches.csv  profiles/rNN_launches.md
  python profiles/extract_ncu.py full     gpurun_out/prof_x.ncu-rep [...] profiles/rNN_ncu_full.md
  python profiles/extract_ncu.py traffic  gpurun_out/prof_x.ncu-rep [...] profiles/ncu_traffic.json

`launches` reads the CSV of an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` pass and
prints per-kernel launch counts, mean/min/max duration and share of the total GPU time.  `full` reads
`--set full` reports with `ncu -i … --page raw --csv` and keeps the metrics DESIGN.md argues from.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
    "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = name.replace("kwg::", "").replace("void ", "")
    return name.split("(")[0]


def launches(path, out):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        agg.setdefault(k, []).append(float(r["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list summary (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("source: %s, %d launches, %.3f ms of GPU time (cold-cache, serialised: shares matter, not absolutes)\n\n" %
                (path, sum(len(v) for v in agg.values()), total / 1e6))
        f.write("| kernel | grid | block | launches | mean us | min us | max us | share |\n|---|---|---|---|---|---|---|---|\n")
        for (k, g, b), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f | %.3f |\n" %
                    (k, g, b, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3, sum(v) / total))
    print("wrote", out)


def full(reps, out):
    with open(out, "w") as f:
        f.write("# ncu --set full summaries (per launch; --clock-control none)\n")
        for rep in reps:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(txt)))
            hdr, units = rows[0], rows[1]
            for r in rows[2:]:
                d = dict(zip(hdr, r))
                f.write("\n## %s  (%s)\n\n" % (short(d["Kernel Name"]), rep.split("/")[-1]))
                f.write("kernel: `%s`  grid %s block %s\n\n| metric | value | unit |\n|---|---|---|\n" %
                        (d["Kernel Name"], d["Grid Size"], d["Block Size"]))
                for m in KEEP:
                    if m in d:
                        f.write("| %s | %s | %s |\n" % (m, d[m], units[hdr.index(m)]))
                try:
                    rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"])
                    ur = units[hdr.index("dram__bytes_read.sum")]; uw = units[hdr.index("dram__bytes_write.sum")]
                    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                    tot = rd * sc[ur] + wr * sc[uw]
                    t = float(d["gpu__time_duration.sum"]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
                    f.write("| **traffic (dram read+write)** | %.4g | byte |\n| **dram GB/s under ncu** | %.1f | GB/s |\n" % (tot, tot / t / 1e9))
                except (KeyError, ValueError):
                    pass
    print("wrote", out)


def traffic(reps, out):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel -> the file bench.py reads roofline.traffic from,
    stamped with the commit the captures belong to (HEAD at extraction time; capture and extract from the same tree)."""
    import json
    import os
    sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    kernels = {}
    try:
        with open(out) as f:
            kernels = json.load(f).get("kernels", {})       # other kernels' entries from earlier captures stay
    except Exception:
        pass
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = short(d["Kernel Name"]).split("<")[0]
            tot = float(d["dram__bytes_read.sum"]) * sc[units[hdr.index("dram__bytes_read.sum")]] + \
                float(d["dram__bytes_write.sum"]) * sc[units[hdr.index("dram__bytes_write.sum")]]
            kernels[name] = tot
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
    with open(out, "w") as f:
        json.dump({"commit": head, "reports": [os.path.basename(r) for r in reps], "unit": "bytes per launch (dram read + write)",
                   "kernels": kernels}, f, indent=1)
        f.write("\n")
    print("wrote", out, kernels)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2:-1], sys.argv[-1])
    else:
        full(sys.argv[2:-1], sys.argv[-1])
