import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster
N = 1_000_000
r, vr, vt, m = models.plummer(N, seed=77)
c = Cluster(N, n_star_dim=N + 64)
c.load(r, vr, vt, m * N)
c.calc_potential_new(); c.compute_energy(); c.upload()
c.profile(True)
rows = []
for step in range(120):
    c.profile_read(reset=True)
    c.evolve(t_end=1e9, max_steps=1, seed=11, tidally_strip=0, GAMMA=0.1)
    inner = c.profile_read(reset=True)["sigma"][0]
    a, sg = c.dev_calc_sigma_r(20)
    st = c.sigma_stats
    n = c.clus.N_MAX
    rows.append((step, n, (n - 1) % 2048, round(inner, 2), st["rerun"], st["hard_tiles"], st["broken"]))
slow = [x for x in rows if x[3] > 2.0]
print("steps with a slow internal calc_sigma_r:", len(slow), "of", len(rows))
for x in rows[:12] + slow[:12]: print(x)
