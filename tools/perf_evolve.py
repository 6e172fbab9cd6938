"""Where a timestep of Cluster.evolve goes: stage timers of the library + wall clock per call.  python tools/perf_evolve.py [N] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
r, vr, vt, m = models.plummer(N, seed=77)
c = Cluster(N, n_star_dim=N + 64)
c.load(r, vr, vt, m * N)
c.calc_potential_new()
c.compute_energy()
c.upload()
c.evolve(t_end=1e9, max_steps=3, seed=11, tidally_strip=0, GAMMA=0.1)
for variant in (0, 1, None):
    c.profile(True); c.profile_read(reset=True); c.sync()
    t0 = time.time()
    c.evolve(t_end=1e9, max_steps=steps, seed=11, tidally_strip=variant, GAMMA=0.1)
    c.sync()
    wall = (time.time() - t0) / steps * 1e3
    pr = c.profile_read()
    c.profile(False)
    print(f"N={N} tidally_strip={variant}: {wall:.2f} ms wall per timestep; device stages (ms per step): " +
          " ".join(f"{k}={v[0] / steps:.2f}" for k, v in pr.items() if v[0] > 0) + f"; sum {sum(v[0] for v in pr.values()) / steps:.2f}", flush=True)
# wall clock per call of one timestep
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
c.evolve(t_end=1e9, max_steps=steps, seed=11, tidally_strip=0, GAMMA=0.1)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
