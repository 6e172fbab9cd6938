#!/usr/bin/env python
"""Wall-clock of the host-buffer entry points (pinned arrays) for a few settings of the pipeline knobs.  e2e_probe.py N "ENV=..," ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster
N = int(float(sys.argv[1]))
r, vr, vt, m = models.plummer(N, seed=1234)
for cfg in sys.argv[2:] or [""]:
    for k in [k for k in os.environ if k.startswith("CMCB200_")]:
        os.environ.pop(k)
    os.environ.update(dict(kv.split("=") for kv in cfg.split(",") if kv))
    c = Cluster(N, n_star_dim=N + 4096, pinned=True)
    c.load(r, vr, vt, np.ones(N)); c.bin_slots_in_use = 0
    c.calc_potential_new(); c.compute_energy()
    ts = []
    for it in range(4):
        t0 = time.perf_counter(); c.new_orbits_calculate(seed=it, want_status=False); t1 = time.perf_counter()
        ts.append(t1 - t0)
    t0 = time.perf_counter(); c.qsorts_new(); t1 = time.perf_counter(); c.calc_potential_new(); t2 = time.perf_counter()
    print(f"[{cfg or 'defaults'}] new_orbits_calculate {min(ts[1:]) * 1e3:.1f} ms (first {ts[0] * 1e3:.1f})  qsorts_new {1e3 * (t1 - t0):.1f} ms  calc_potential_new {1e3 * (t2 - t1):.1f} ms", flush=True)
    c.close()
