#!/usr/bin/env python
"""Times the stages of the resident step (CUDA events through the library's stage profiler) for one or more settings of the tuning knobs.
The first step starts from a fresh Plummer state; the later ones contain stars with v_r == 0 whose linear scans are queued.

    perf_orbit.py N [steps] ["CMCB200_ORBIT_LADDERS=0,CMCB200_ORBIT_CTAS=3" ...]     one context per knob string ("" = defaults)
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
configs = sys.argv[3:] or [""]
r, vr, vt, m = models.plummer(N, seed=1234)
for cfg in configs:
    knobs = dict(kv.split("=") for kv in cfg.split(",") if kv)
    for k in [k for k in os.environ if k.startswith("CMCB200_")]:
        os.environ.pop(k, None)
    os.environ.update(knobs)
    c = Cluster(N, n_star_dim=N + 4096)
    c.load(r, vr, vt, np.ones(N))
    c.calc_potential_new()
    c.compute_energy()
    c.upload()
    c.profile(True)
    for s in range(steps):
        c.profile_read(reset=True)
        res = c.dev_step(seed=1)
        p = c.profile_read(reset=True)
        orbit = p['orbit_prepare'][0] + p['orbit_scan'][0] + p['orbit_tries'][0] + p['orbit_tables'][0]
        total = sum(v[0] for v in p.values())
        print(f"N={N} [{cfg or 'defaults'}] step {s}: total {total:.3f} ms orbit {orbit:.3f} ms  deferred={res.n_deferred} exact={res.n_exact} "
              f"brent={res.n_brent} status={list(res.n_status)} tries={(res.n_draws / max(res.n_status[0], 1) - 1) / 2:.3f} | "
              + " ".join(f"{k}={v[0]:.3f}" for k, v in p.items() if v[0] > 0), flush=True)
    c.close()
