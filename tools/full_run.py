"""A whole CMC run on the resident arrays: a Plummer sphere evolved by two-body relaxation to (near) core collapse with the configuration of the
reference's examples/PlummerSphere.ini, every N-sized operation of the main loop on the GPU (Cluster.evolve), and its Lagrange radii next to
those of the reference's own example run (tests/golden/plummer_lagrad_run.npz, from docs/sphinx/source/example_output/plummer.lagrad.dat).
python tools/full_run.py [N] [t_end]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from cmc_cosmic_b200 import models
from cmc_cosmic_b200.cluster import Cluster

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
t_end = float(sys.argv[2]) if len(sys.argv) > 2 else 1.40
ref = np.load(os.path.join(ROOT, "tests", "golden", "plummer_lagrad_run.npz"))
fr, t_ref, r_ref = ref["mass_pc"], ref["t"], ref["r"]
r, vr, vt, m = models.plummer(N, seed=77)
c = Cluster(N, n_star_dim=N + 64)
c.load(r, vr, vt, m * N)
c.calc_potential_new()
c.compute_energy()
c.upload()
c.transfer_bytes(reset=True)
c.sync()
t0 = time.time()
hist = c.evolve(t_end=t_end, max_steps=30000, seed=11, THETASEMAX=1.0, tidally_strip=0, GAMMA=0.1, mass_pc=fr, record_every=5, min_dt=2e-5)
c.sync()
wall = time.time() - t0
steps = len(hist) * 5
up, down = c.transfer_bytes()
print(f"N = {N}: {steps} timesteps to t = {hist[-1][0]:.4f} in {wall:.2f} s wall ({wall / steps * 1e3:.3f} ms per timestep incl. the Lagrange radii every 5 steps); "
      f"N_MAX {N} -> {hist[-1][2]}; bytes over PCIe during the run: {up} up, {down} down (the reference run: {int(ref['n_steps'])} steps to t = {float(ref['t_end']):.4f})")
th = np.array([h[0] for h in hist])
lag = np.array([h[3] for h in hist])
sel = [int(np.argmin(np.abs(fr - f))) for f in (0.001, 0.01, 0.05, 0.1, 0.3, 0.5, 0.7, 0.9)]
print("Lagrange radii, this run / reference run, at mass fractions " + " ".join(f"{fr[j]:g}" for j in sel))
for tq in np.arange(0.0, t_end + 1e-9, 0.1):
    w = np.abs(th - tq) <= 0.015
    if not w.any():
        continue
    mine = lag[w].mean(axis=0)
    want = r_ref[int(np.argmin(np.abs(t_ref - tq)))]
    print(f"t = {tq:4.2f}  " + "  ".join(f"{mine[j]:.4f}/{want[j]:.4f} ({(mine[j] / want[j] - 1) * 100:+5.1f}%)" for j in sel))
