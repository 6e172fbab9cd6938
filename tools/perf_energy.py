#!/usr/bin/env python
"""Times the energy scheme's device passes inside full timesteps on the resident arrays, in the reference's order
(src/cmc/cmc.c:302-353): energy_conservation1 -> orbit step -> energy_conservation2 -> ComputeIntermediateEnergy (EI + transfer) -> sort ->
potential -> energy_conservation3 (set_velocities3) -> ComputeEnergy.  CUDA events through the library's stage profiler.

    perf_energy.py N [steps]
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
r, vr, vt, m = models.plummer(N, seed=1234)
c = Cluster(N, n_star_dim=N + 4096)
c.load(r, vr, vt, np.ones(N))
c.calc_potential_new()
c.compute_energy()
c.upload()
c.profile(True)


def timed(f):
    c.profile_read(reset=True)
    out = f()
    p = c.profile_read(reset=True)
    return out, sum(v[0] for v in p.values())


for s in range(steps):
    _, t1 = timed(c.dev_energy_conservation1)
    _, to = timed(lambda: c.dev_new_orbits(seed=1))
    _, t2 = timed(c.dev_energy_conservation2)
    import time
    w0 = time.perf_counter()
    _, tt = timed(lambda: c.dev_tidally_strip_stars(1))
    tw = (time.perf_counter() - w0) * 1e3
    _, ti = timed(c.dev_compute_intermediate_energy)
    _, ts = timed(c.dev_qsorts_new)
    _, tp = timed(c.dev_calc_potential_new)
    left, t3 = timed(c.dev_energy_conservation3)
    _, te = timed(c.dev_compute_energy)
    c.tcount += 1
    print(f"N={N} step {s}: energy_conservation1 {t1:.3f} ms | orbit {to:.3f} | energy_conservation2 {t2:.3f} | ComputeIntermediateEnergy {ti:.3f} | "
          f"tidally_strip_stars {tt:.3f} (wall {tw:.2f}) | sort {ts:.3f} | potential {tp:.3f} | energy_conservation3 {t3:.3f} (leftover Eexcess {left:.3e}) | ComputeEnergy {te:.3f}", flush=True)
for s in range(steps):
    _, tt = timed(lambda: c.dev_timestep(seed=1))
    print(f"N={N} fused timestep {s} (Cluster.dev_timestep: one lookup for energy_conservation2 + ComputeIntermediateEnergy, transfer inside the gather): {tt:.3f} ms", flush=True)
c.close()
