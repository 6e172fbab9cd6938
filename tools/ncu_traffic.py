#!/usr/bin/env python
"""profiles/traffic_rNN.json and a readable per-kernel summary from an .ncu-rep (read here with `ncu -i`, no GPU needed).
    ncu_traffic.py capture.ncu-rep N_STARS out.json out_summary.txt"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instructions")]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(rep, n_stars, out_json, out_txt):
    n_stars = float(n_stars)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    traffic, lines, seen = {}, [], {}
    for d in data:
        full = d[col["Kernel Name"]]
        short = re.sub(r"^void ", "", full).split("(")[0]
        base = short.split("<")[0]
        seen[short] = seen.get(short, 0) + 1
        rd = to_bytes(d[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(d[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        lines.append(f"== {short}  (launch #{seen[short]} of this kernel in the capture)")
        for k, label in KEYS:
            if k in col:
                lines.append(f"   {label:22s} {d[col[k]]:>18s} {units[col[k]]}")
        lines.append(f"   {'dram bytes per star':22s} {(rd + wr) / n_stars:18.1f} (read {rd / n_stars:.1f} + write {wr / n_stars:.1f}; N = {n_stars:.0f})")
        if base not in traffic:  # the first launch of each kernel works on all N stars
            traffic[base] = {"dram_bytes_per_star": round((rd + wr) / n_stars, 2), "dram_read_per_star": round(rd / n_stars, 2),
                             "dram_write_per_star": round(wr / n_stars, 2), "kernel": short, "capture": rep.split("/")[-1],
                             "time_ms": float(d[col["gpu__time_duration.sum"]]) if units[col["gpu__time_duration.sum"]] == "ms" else None}
    json.dump(traffic, open(out_json, "w"), indent=1)
    open(out_txt, "w").write("\n".join(lines) + "\n")
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
