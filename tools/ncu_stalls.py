#!/usr/bin/env python
"""Stall-sample breakdown of one kernel of an .ncu-rep captured with --import-source on (read here with `ncu -i`, no GPU needed): the share
of every stall reason, the samples between consecutive barriers (the phases of a tile), and the hottest instructions.
    ncu_stalls.py capture.ncu-rep KERNEL_ID [n_top]"""
import csv
import subprocess
import sys


def main(rep, kid, n_top=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    print("kernel:", r[0][1][:140])
    h = r[1]
    rows = [x for x in r[2:] if len(x) >= len(h) - 2]
    ix = {n: i for i, n in enumerate(h)}
    first = [x[ix["Source"]] for x in rows[:3]]
    for k in range(100, len(rows) - 3):                 # the page lists the function twice: keep the first listing
        if [x[ix["Source"]] for x in rows[k:k + 3]] == first:
            rows = rows[:k]
            break

    def g(x, n):
        try:
            return int(x[ix[n]] or 0)
        except (ValueError, IndexError):
            return 0
    tot = sum(g(x, "# Samples") for x in rows)
    print(f"stall samples: {tot} over {len(rows)} instructions")
    reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    for n, v in sorted(((n, sum(g(x, n) for x in rows)) for n in reasons), key=lambda t: -t[1])[:8]:
        print(f"   {n:26s} {v:7d}  {100.0 * v / tot:5.1f} %")
    print("samples between barriers (phases of one tile / iteration):")
    acc = last = 0
    for i, x in enumerate(rows):
        acc += g(x, "# Samples")
        src = x[ix["Source"]]
        if "BAR.SYNC" in src or i == len(rows) - 1:
            print(f"   up to instruction {i:5d} ({src.strip()[:28]:28s}) {acc - last:7d}  {100.0 * (acc - last) / tot:5.1f} %")
            last = acc
    print("hottest instructions:")
    for i in sorted(sorted(range(len(rows)), key=lambda i: -g(rows[i], "# Samples"))[:int(n_top)]):
        print(f"   {i:5d} {g(rows[i], '# Samples'):6d}  {rows[i][ix['Source']].strip()[:100]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
