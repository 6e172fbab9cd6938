"""Times calc_sigma_r and dynamics_apply's pair loop on the resident arrays (stage timers of the library).  python tools/perf_sigma.py [N] [imf|equal]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
masses = sys.argv[2] if len(sys.argv) > 2 else "equal"
r, vr, vt, m = models.plummer(N, seed=3)
if masses == "imf":
    m = models.kroupa_like_masses(N, seed=4)
c = Cluster(N, n_star_dim=N + 64)
c.load(r, vr, vt, m * N)
c.upload()
c.dev_calc_potential_new()
c.dev_compute_energy()
c.profile(True)
for rep in range(4):
    c.profile_read(reset=True)
    t0 = time.time()
    a, sg = c.dev_calc_sigma_r(20)
    t1 = time.time()
    res = c.dev_dynamics_apply(1e-3, seed=rep)
    c.sync()
    t2 = time.time()
    pr = c.profile_read()
    print(f"N={N} {masses} rep {rep}: calc_sigma_r {pr['sigma'][0]:.3f} ms on the device ({(t1 - t0) * 1e3:.1f} ms with its 16 B/star read-back), stats {c.sigma_stats}; "
          f"dynamics_apply {pr['dynamics'][0]:.3f} ms ({(t2 - t1) * 1e3:.1f} ms wall), Nrel {res.Nrel}, encounters {res.n_encounters}", flush=True)
