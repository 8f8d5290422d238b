#!/usr/bin/env python
"""Reads the round's `ncu --set full` captures (gpurun_out/r2_prof_<config>_<format>[_f32].ncu-rep, scripts/r2_ncu.sh) here on
the CPU box with `ncu -i ... --page raw --csv` and writes
  profiles/ncu_traffic.json   {"c5/crs": DRAM bytes of ONE multiply's launches of the dominant kernel, ...}   (bench.py's traffic)
  profiles/r2_ncu_kernels.md  one row per captured launch: duration, DRAM read/write, DRAM and L2 throughput, hit rates, regs
"""
import csv
import glob
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"gpu__time_duration.sum": "dur", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "dram__bytes.sum.per_second": "dram_bps", "lts__t_sector_hit_rate.pct": "l2_hit",
        "l1tex__t_sector_hit_rate.pct": "l1_hit", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps",
        "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__inst_executed_pipe_uniform.sum": "uni"}
UNIT = {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    start = out.find('"ID"')
    if start < 0:
        return []
    rd = list(csv.reader(io.StringIO(out[start:])))
    head, units, body = rd[0], rd[1], rd[2:]
    res = []
    for r in body:
        d = {"kernel": r[head.index("Kernel Name")]}
        for i, h in enumerate(head):
            if h in WANT and i < len(r):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[WANT[h]] = v * UNIT.get(units[i], 1.0)
        res.append(d)
    return res


def main():
    traffic, lines = {}, ["# Round 2: `ncu --set full --clock-control none` captures of the dominant kernels (one B200)\n\n",
                          "Produced by `scripts/r2_ncu.sh` on the GPU box, read here with `scripts/ncu_traffic.py`.  Durations under ncu are\n",
                          "cold-cache and serialised: compare shares and byte counts, not times.  alg = compulsory bytes of the launch.\n\n",
                          "| capture | kernel | launch | duration | DRAM read | DRAM write | DRAM GB/s | DRAM % of peak | L2 % | L2 hit % | L1 hit % | regs | grid x block |\n",
                          "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n"]
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2_prof_*.ncu-rep"))):
        tag = os.path.basename(rep)[len("r2_prof_"):-len(".ncu-rep")]
        rows = rows_of(rep)
        if not rows:
            continue
        total = 0.0
        for i, d in enumerate(rows):
            total += d.get("rd", 0) + d.get("wr", 0)
            lines.append("| %s | `%s` | %d | %.1f us | %.3f GB | %.3f GB | %.0f | %.1f | %.1f | %.1f | %.1f | %d | %d x %d |\n" % (
                tag, re.sub(r"\(.*", "", d["kernel"])[:60], i, d.get("dur", 0) * 1e6, d.get("rd", 0) / 1e9, d.get("wr", 0) / 1e9,
                d.get("dram_bps", 0) / 1e9, d.get("dram_pct", 0), d.get("l2_pct", 0), d.get("l2_hit", 0), d.get("l1_hit", 0), int(d.get("regs", 0)),
                int(d.get("grid", 0)), int(d.get("block", 0))))
        m = re.match(r"(c\d)_([a-z0-9]+)(_f32)?$", tag)
        if m:
            key = "%s/%s%s" % (m.group(1), m.group(2), "_f32" if m.group(3) else "")
            traffic[key] = int(total)                              # one multiply: the capture holds exactly one step's launches
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    open(os.path.join(ROOT, "profiles", "r2_ncu_kernels.md"), "w").write("".join(lines))
    print("".join(lines))
    print(traffic)


if __name__ == "__main__":
    main()
