#!/usr/bin/env python
"""Kernel shares of the timed steps from an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --steps 2 --warmup 1`.
    launch_shares.py launches.csv [bench.json]  ->  stdout (profiles/launch_shares_rNN.txt)
The steps are found as the repetitions of the launch sequence that starts with orbit_stats_reset / orbit_pack_tables of a resident step; the LAST
two complete repetitions before the e2e part are the timed ones.  ncu serialises the launches and runs them cold: the SHARES are comparable with
the bench's own CUDA-event stage timers, not the absolute times."""
import collections
import csv
import json
import re
import sys


def main(path, bench=None):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows:
        if len(r) < len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        t = float(r[col["Metric Value"]]) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}[r[col["Metric Unit"]]]
        name = re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0].split("<")[0]
        launches.append((name, t))
    # a resident step = from one orbit_pack_tables to the launch before the next one that follows a potential rebuild (pot_phi0)
    starts = [i for i, (n, _) in enumerate(launches) if n == "orbit_pack_tables"]
    ends = [i for i, (n, _) in enumerate(launches) if n == "pot_phi0"]
    steps = []
    for s in starts:
        e = next((x for x in ends if x > s), None)
        nxt = next((x for x in starts if x > s), len(launches))
        if e is not None and e < nxt:
            steps.append((s, e))
    print(f"{len(launches)} launches in the capture, {len(steps)} complete orbit -> sort -> potential sequences")
    for k, (s, e) in enumerate(steps):
        acc = collections.OrderedDict()
        for n, t in launches[s:e + 1]:
            acc[n] = acc.get(n, 0.0) + t
        tot = sum(acc.values())
        print(f"sequence {k} (launches {s}-{e}): {e - s + 1} launches, total {tot:.1f} us")
        for n, t in sorted(acc.items(), key=lambda x: -x[1])[:12]:
            print(f"  {n:36s} {t:9.1f} us  {100 * t / tot:5.1f} %")
    if bench:
        d = json.loads([l for l in open(bench).read().splitlines() if l.startswith("{")][-1])
        st = d["stage_ms_per_step"]
        print("CUDA-event stage timers of the bench run proper (" + bench + "), ms per step: " + ", ".join(f"{k} {v}" for k, v in st.items() if v > 0))
        print(f"  ms_per_step {d['ms_per_step']:.3f}: " + ", ".join(f"{k} {100 * v / d['ms_per_step']:.1f} %" for k, v in st.items() if v > 0.3))


if __name__ == "__main__":
    main(*sys.argv[1:3])
