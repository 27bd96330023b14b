#!/usr/bin/env python3
"""Turns the ncu artefacts brought back in gpurun_out/ into the small text/JSON summaries that are
committed under profiles/ (gpurun_out/ is scratch).  Usage:
  python profiles/summarize.py launches gpurun_out/launches3.csv profiles/r01_launches_flat_b1.txt
  python profiles/summarize.py full gpurun_out/prof_scan.ncu-rep profiles/r01_flat_scan_full.txt [traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u in ("nsecond", "ns") else v * 1000 if u in ("msecond", "ms") else v
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as o:
        o.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        o.write(f"# source: {src}; total {tot/1000:.3f} ms over {sum(len(v) for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            o.write(f"{sum(v)/tot*100:6.2f}%  n={len(v):5d}  avg={sum(v)/len(v):10.2f} us  min={min(v):10.2f}  max={max(v):10.2f}  {k[:140]}\n")
    pass


def full(src, dst, traffic=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    recs = []
    with open(dst, "w") as o:
        o.write(f"# ncu --set full --clock-control none; source: {src}\n")
        for row in r[2:]:
            name = row[hdr.index("Kernel Name")]
            o.write(f"\n== {name[:160]}\n")
            rec = {}
            for i, h in enumerate(hdr):
                if h in KEEP:
                    o.write(f"  {h:90s} {row[i]:>18s} {units[i]}\n")
                    rec[h] = (row[i], units[i])
            recs.append(rec)
    pass
    if traffic and recs:
        def gb(x):
            v, u = x
            v = float(v.replace(",", ""))
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        t = [gb(x["dram__bytes_read.sum"]) + gb(x["dram__bytes_write.sum"]) for x in recs]
        json.dump({"dram_bytes_per_launch": sum(t) / len(t), "launches": len(t), "source": src}, open(traffic, "w"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
