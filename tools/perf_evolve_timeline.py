"""Stage timers along a whole run: where do the timesteps of Cluster.evolve go as the cluster evolves?  python tools/perf_evolve_timeline.py [N] [t_end] [chunk]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
t_end = float(sys.argv[2]) if len(sys.argv) > 2 else 1.4
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 200
r, vr, vt, m = models.plummer(N, seed=77)
c = Cluster(N, n_star_dim=N + 64)
c.load(r, vr, vt, m * N)
c.calc_potential_new()
c.compute_energy()
c.upload()
c.profile(True)
total = 0
while getattr(c, "TotalTime", 0.0) < t_end and total < 40000:
    c.profile_read(reset=True); c.sync()
    t0 = time.time()
    h = c.evolve(t_end=t_end, max_steps=chunk, seed=11, tidally_strip=0, GAMMA=0.1)
    c.sync()
    n = len(h)
    total += n
    wall = (time.time() - t0) / max(n, 1) * 1e3
    pr = c.profile_read()
    st = c.sigma_stats if hasattr(c, "sigma_stats") else {}
    print(f"steps {total - n}..{total} t={c.TotalTime:.4f} Dt={c.Dt:.2e} N_MAX={c.clus.N_MAX}: {wall:.2f} ms wall/step; " +
          " ".join(f"{k}={v[0] / max(n, 1):.2f}" for k, v in pr.items() if v[0] / max(n, 1) > 0.05) + f" | orbit: deferred={c.last_orbit.n_deferred} exact={c.last_orbit.n_exact} "
          f"draws/star={c.last_orbit.n_draws / max(c.clus.N_MAX, 1):.2f} status={list(c.last_orbit.n_status)}", flush=True)
